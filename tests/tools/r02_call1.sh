#!/bin/bash
# round 2, first GPU call: state of the tree as round 1 left it + the switches written blind
set -u
export CLIPK_BENCH_QUICK=1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
echo "=== gpu suite"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "=== pending distill tests"; CLIPK_TEST_PENDING=1 timeout 200 python -m pytest tests/test_distill_gpu.py -q 2>&1 | tail -15
run() { local label=$1; shift; echo "=== $label"; env "$@" timeout 150 python bench.py --steps 20 --warmup 5 2>gpurun_out/err_$$.log | python -c "
import json,sys
j=json.loads(sys.stdin.readline()); r=j['roofline']['breakdown_ms']
print({'ms_per_step': round(j['ms_per_step'],4), 'e2e_ms': round(j['e2e']['ms_per_step'],4), 'serial': round(j['e2e']['serial_ms_per_step'],4), 'clk': j['clocks']['sm_mhz'], **{k: round(v,4) for k,v in r.items()}})" || tail -5 gpurun_out/err_$$.log; }
run "baseline" CLIPK_BWD_STREAMS=1
run "two-stream backward" CLIPK_BWD_STREAMS=2
run "panel 96MB" CLIPK_PANEL_MB=96
run "panel 96MB two-stream" CLIPK_PANEL_MB=96 CLIPK_BWD_STREAMS=2
run "panel 64MB two-stream" CLIPK_PANEL_MB=64 CLIPK_BWD_STREAMS=2
echo "=== parity with the two-stream backward"
CLIPK_BWD_STREAMS=2 timeout 300 python -m pytest tests/test_parity_gpu.py -q -x 2>&1 | tail -5
