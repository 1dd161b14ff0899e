#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== gpu suite"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "=== shard timing 8192 x 32768"; CLIPK_VERBOSE=1 SHARD_TIME=1 timeout 100 python tests/tools/shard_step.py 8192 32768 512 1 2>&1 | grep -E "timing|panel" | sort -u | tail -2
echo "=== step timeline, 8-GPU-shard-sized single-GPU problem (4096 rows) and N = 32768"
timeout 100 python tests/tools/step_timeline.py 4096 512 2 2>&1 | tail -45
timeout 100 python tests/tools/step_timeline.py 32768 512 2 2>&1 | grep -v "grad_sweep\|gemm_pair" | tail -40
