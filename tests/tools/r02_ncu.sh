#!/bin/bash
# ncu evidence (one GPU, final defaults): launch list of one rank's kernels at the one-GPU shape and at an 8-GPU shard
# shape, full sets of the tensor-core kernels.  Every ncu run follows a plain run of the same command that exited 0.
set -u
mkdir -p gpurun_out
for shape in "32768 32768 512" "4096 32768 512"; do
  tag=$(echo $shape | tr ' ' 'x')
  python tests/tools/shard_step.py $shape 2 > gpurun_out/plain_$tag.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_zz_$tag.csv \
      python tests/tools/shard_step.py $shape 2 > gpurun_out/ncu_l_$tag.log 2>&1
  python tests/tools/ncu_summary.py gpurun_out/launches_zz_$tag.csv 2>/dev/null | head -8
done
python tests/tools/shard_step.py 32768 32768 512 1 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"fwd_sweep_kernel|grad_sweep_kernel|gemm_pair_kernel" -c 4 \
    -o gpurun_out/prof_r02zz python tests/tools/shard_step.py 32768 32768 512 1 > gpurun_out/ncu_f.log 2>&1
ls -la gpurun_out/prof_r02zz.ncu-rep
