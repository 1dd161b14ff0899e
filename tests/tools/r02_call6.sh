#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== parity tests with the swizzled recompute staging"; timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_step_gpu.py -q -x 2>&1 | tail -4
echo "=== shard timing"; SHARD_TIME=1 timeout 100 python tests/tools/shard_step.py 4096 32768 512 1 2>&1 | tail -1
echo "=== bench N=1 quick"; CLIPK_BENCH_QUICK=1 timeout 300 python bench.py --skip-extras > gpurun_out/bench_r2h_n1.json 2> gpurun_out/bench_r2h_n1.err || tail -20 gpurun_out/bench_r2h_n1.err
python tests/tools/show_bench.py gpurun_out/bench_r2h_n1.json 2>/dev/null | head -14
echo "=== ncu of the recompute kernel"
python tests/tools/shard_step.py 32768 32768 512 1 > gpurun_out/plain6.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"grad_sweep_kernel" -c 1 -o gpurun_out/prof_r02h_grad python tests/tools/shard_step.py 32768 32768 512 1 > gpurun_out/ncu6.log 2>&1
ls -la gpurun_out/prof_r02h_grad.ncu-rep
