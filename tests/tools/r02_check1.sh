#!/bin/bash
# one GPU: the -m gpu suite, then the default bench line (final defaults of round 2)
set -u
mkdir -p gpurun_out
echo "=== gpu suite"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "=== bench N=1 (default flags)"
timeout 600 python bench.py > gpurun_out/bench_r2zz_n1.json 2> gpurun_out/bench_r2zz_n1.err || tail -20 gpurun_out/bench_r2zz_n1.err
python tests/tools/show_bench.py gpurun_out/bench_r2zz_n1.json
