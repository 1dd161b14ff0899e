#!/bin/bash
# round 2, second GPU call: the fused step on one GPU
set -u
export CLIPK_BENCH_QUICK=1
mkdir -p gpurun_out
echo "=== step tests"; timeout 600 python -m pytest tests/test_step_gpu.py -x -q 2>&1 | tail -15
echo "=== gpu suite"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15
echo "=== bench (old harness)"; timeout 200 python bench.py --steps 20 --warmup 5 2>gpurun_out/err2.log | python -c "
import json,sys
j=json.loads(sys.stdin.readline()); r=j['roofline']['breakdown_ms']
print({'ms_per_step': round(j['ms_per_step'],4), 'e2e_ms': round(j['e2e']['ms_per_step'],4), 'serial': round(j['e2e']['serial_ms_per_step'],4), 'launches': j['gpu_launches'], 'clk': j['clocks']['sm_mhz'], **{k: round(v,4) for k,v in r.items()}})" || tail -20 gpurun_out/err2.log
