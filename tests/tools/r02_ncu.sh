#!/bin/bash
# ncu evidence of round 2 (one GPU): launch lists at the one-GPU shape and at an 8-GPU shard shape, full sets of the three
# tensor-core kernels at both.  Every ncu run follows a plain run of the same command that exited 0.
set -u
mkdir -p gpurun_out
for shape in "32768 32768 512" "4096 32768 512"; do
  tag=$(echo $shape | tr ' ' 'x')
  python tests/tools/shard_step.py $shape 2 > gpurun_out/plain_$tag.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$tag.csv \
      python tests/tools/shard_step.py $shape 2 > gpurun_out/ncu_l_$tag.log 2>&1
  tail -2 gpurun_out/plain_$tag.log
done
for shape in "32768 32768 512" "4096 32768 512"; do
  tag=$(echo $shape | tr ' ' 'x')
  python tests/tools/shard_step.py $shape 1 > gpurun_out/plain2_$tag.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"fwd_sweep_kernel|grad_sweep_kernel|gemm_pair_kernel|prep_kernel" -c 6 \
      -o gpurun_out/prof_r02_$tag python tests/tools/shard_step.py $shape 1 > gpurun_out/ncu_f_$tag.log 2>&1
  ls -la gpurun_out/prof_r02_$tag.ncu-rep
done
