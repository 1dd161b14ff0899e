#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== new tests"; timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_ict_gpu.py tests/test_step_gpu.py -q -k "split_row_col or ict or logits or step" 2>&1 | tail -15
echo "=== bench N=1 quick"; CLIPK_BENCH_QUICK=1 timeout 300 python bench.py --skip-extras > gpurun_out/bench_r2g_n1.json 2> gpurun_out/bench_r2g_n1.err || tail -20 gpurun_out/bench_r2g_n1.err
python tests/tools/show_bench.py gpurun_out/bench_r2g_n1.json | head -14
