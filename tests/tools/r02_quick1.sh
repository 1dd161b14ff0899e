#!/bin/bash
# one GPU: step + parity tests, then the default bench without extras: bash tests/tools/r02_quick1.sh [tag]
set -u
tag=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_step_gpu.py tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do
CLIPK_BENCH_QUICK=1 timeout 300 python bench.py --skip-extras 2>gpurun_out/bench_$tag.err | tee gpurun_out/bench_$tag.json | python -c "
import json,sys
j=json.loads(sys.stdin.readline()); print(round(j['ms_per_step'],4), 'e2e', round(j['e2e']['ms_per_step'],3), 'parity', j['parity'].get('ok'), {k[:10]: (round(v['launches_per_step']), round(v['ms_per_step'],3)) for k,v in j['roofline']['kernels'].items() if v['ms_per_step']>0.02})"
done
