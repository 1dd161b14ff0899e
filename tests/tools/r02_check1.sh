#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== gpu suite"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "=== bench N=1 quick"; CLIPK_BENCH_QUICK=1 timeout 300 python bench.py --skip-extras > gpurun_out/bench_r2j_n1.json 2> gpurun_out/bench_r2j_n1.err || tail -20 gpurun_out/bench_r2j_n1.err
python tests/tools/show_bench.py gpurun_out/bench_r2j_n1.json 2>/dev/null | head -14
