#!/bin/bash
# panel budget A/B on the one-GPU step (N = 32768, d = 512): bash tests/tools/r02_panel_ab.sh
set -u
for mb in 192 384 512 700 192 512; do echo "one GPU, CLIPK_PANEL_MB=$mb"; CLIPK_PANEL_MB=$mb CLIPK_BENCH_QUICK=1 timeout 300 python bench.py --skip-extras 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.readline()); print(round(j['ms_per_step'],4), 'e2e', round(j['e2e']['ms_per_step'],3), {k[:10]: (round(v['launches_per_step']), round(v['ms_per_step'],3)) for k,v in j['roofline']['kernels'].items() if v['ms_per_step']>0.3})"; done
