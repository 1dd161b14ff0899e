#!/bin/bash
# panel budget A/B at the shard shapes of 2 and 4 GPUs (one GPU; backward of one rank, us) and on the one-GPU step
set -u
for shape in "8192 32768 512" "16384 32768 512"; do
  for mb in 192 288 384; do echo "shape $shape CLIPK_PANEL_MB=$mb"; CLIPK_VERBOSE=1 CLIPK_PANEL_MB=$mb SHARD_TIME=1 timeout 100 python tests/tools/shard_step.py $shape 1 2>&1 | grep -E "timing|panel" | sort -u | tail -2; done
done
for mb in 192 288; do echo "one GPU, CLIPK_PANEL_MB=$mb"; CLIPK_PANEL_MB=$mb CLIPK_BENCH_QUICK=1 timeout 300 python bench.py --skip-extras 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.readline()); print(round(j['ms_per_step'],4), {k[:10]: round(v['ms_per_step'],3) for k,v in j['roofline']['kernels'].items() if v['ms_per_step']>0.3})"; done
