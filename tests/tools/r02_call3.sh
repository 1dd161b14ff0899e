#!/bin/bash
# scheduled gradient-GEMM kernel: correctness, then A/B of the schedule variants at the one-GPU and the 8-GPU-shard shape
set -u
export CLIPK_BENCH_QUICK=1
mkdir -p gpurun_out
echo "=== parity + step tests (scheduled pair kernel on)"; timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_step_gpu.py -x -q 2>&1 | tail -5
run() { local label=$1; shift; echo "=== $label"; env "$@" SHARD_TIME=1 timeout 100 python tests/tools/shard_step.py 4096 32768 512 1 2>&1 | tail -1
  env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --skip-extras 2>gpurun_out/err3.log | python -c "
import json,sys
j=json.loads(sys.stdin.readline()); k=j['roofline']['kernels']
print({'ms': round(j['ms_per_step'],4), 'parity': j['parity']['ok'], 'clk': j['clocks']['sm_mhz'], **{n[:14]: round(v['ms_per_step'],3) for n,v in k.items() if v['ms_per_step']>0.1}})" || tail -5 gpurun_out/err3.log; }
run "old pair kernel"            CLIPK_PAIR_SCHED=0
run "sched, dX pieces of 32"     CLIPK_PAIR_SCHED=1 CLIPK_PAIR_PIECE=32
run "sched, dX pieces of 64"     CLIPK_PAIR_SCHED=1 CLIPK_PAIR_PIECE=64
run "sched, no split"            CLIPK_PAIR_SCHED=1 CLIPK_PAIR_PIECE=0
run "sched, dX+dY pieces of 32"  CLIPK_PAIR_SCHED=2 CLIPK_PAIR_PIECE=32
run "sched, dX+dY pieces of 64"  CLIPK_PAIR_SCHED=2 CLIPK_PAIR_PIECE=64
