#!/bin/bash
# same-call A/B of two builds of the library (clipk/libclipk_old.so, clipk/libclipk_new.so) on the one-GPU step
set -u
L=megatron-clip_b200/clipk
mkdir -p gpurun_out
for v in old new old new old new; do cp $L/libclipk_$v.so $L/libclipk.so; echo "build: $v"
CLIPK_BENCH_QUICK=1 timeout 300 python bench.py --skip-extras 2>gpurun_out/bench_ab_$v.err | tee gpurun_out/bench_ab_$v.json | python -c "
import json,sys
j=json.loads(sys.stdin.readline()); print(round(j['ms_per_step'],4), 'e2e', round(j['e2e']['ms_per_step'],3), 'parity', j['parity'].get('ok'), {k[:10]: (round(v['launches_per_step']), round(v['ms_per_step'],3)) for k,v in j['roofline']['kernels'].items() if v['ms_per_step']>0.3})"
done
cp $L/libclipk_new.so $L/libclipk.so
