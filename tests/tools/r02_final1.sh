#!/bin/bash
# one GPU: the whole -m gpu suite, smoke(), and the default bench line (with the CPU legs)
set -u
mkdir -p gpurun_out
echo "=== gpu suite"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
echo "=== bench N=1 (default flags)"
timeout 600 python bench.py > gpurun_out/bench_r2f_n1.json 2> gpurun_out/bench_r2f_n1.err || tail -20 gpurun_out/bench_r2f_n1.err
python tests/tools/show_bench.py gpurun_out/bench_r2f_n1.json
echo "=== reference arm (driver flags)"
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2f_ref.json 2> gpurun_out/bench_r2f_ref.err ) 2>&1 | tail -4
cat gpurun_out/bench_r2f_ref.json | cut -c1-900
