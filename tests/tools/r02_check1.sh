#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== gpu suite"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "=== shard timing (default budget)"; SHARD_TIME=1 timeout 100 python tests/tools/shard_step.py 4096 32768 512 1 2>&1 | tail -1
CLIPK_VERBOSE=1 timeout 100 python tests/tools/shard_step.py 4096 32768 512 1 2>&1 | grep "panel" | head -2
echo "=== bench N=1 quick"; CLIPK_BENCH_QUICK=1 timeout 300 python bench.py --skip-extras > gpurun_out/bench_r2y_n1.json 2> gpurun_out/bench_r2y_n1.err || tail -20 gpurun_out/bench_r2y_n1.err
python tests/tools/show_bench.py gpurun_out/bench_r2y_n1.json 2>/dev/null | head -6
