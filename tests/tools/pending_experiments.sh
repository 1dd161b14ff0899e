#!/bin/bash
# First GPU call of the next session: everything that was written without GPU time, in one go.
#   gpurun --timeout 400 -- 'bash tests/tools/pending_experiments.sh 1 > gpurun_out/pending1.log 2>&1'
#   gpurun --gpus 8 --timeout 300 -- 'bash tests/tools/pending_experiments.sh 8 > gpurun_out/pending8.log 2>&1'
# Every variant is its own process: the switches are read once per process.
set -u
N=${1:-1}
export CLIPK_BENCH_QUICK=1
run() {  # run <label> <env assignments...>: one short bench line
    local label=$1; shift
    echo "=== $label"
    if [ "$N" = 1 ]; then
        env "$@" timeout 120 python bench.py --steps 20 --warmup 5 | python -c "
import json,sys
j=json.loads(sys.stdin.readline()); r=j['roofline']['breakdown_ms']
print({'ms_per_step': round(j['ms_per_step'],4), 'e2e_ms': round(j['e2e']['ms_per_step'],4), 'frac': round(j['roofline']['frac'],4), **{k: round(v,4) for k,v in r.items()}})"
    else
        env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
            --master-port 29517 bench.py --gpus "$N" --steps 20 --warmup 5 | python -c "
import json,sys
j=json.loads(sys.stdin.readline()); r=j['roofline']['breakdown_ms']
print({'ms_per_step': round(j['ms_per_step'],4), 'e2e_ms': round(j['e2e']['ms_per_step'],4), **{k: round(v,4) for k,v in r.items()}})"
    fi
}
if [ "$N" = 1 ]; then
    echo "=== pending GPU tests (clipk/distill.py)"
    CLIPK_TEST_PENDING=1 timeout 120 python -m pytest tests/test_distill_gpu.py -q 2>&1 | tail -15
    run "baseline"                      CLIPK_BWD_STREAMS=1
    run "two-stream backward"           CLIPK_BWD_STREAMS=2
    echo "=== parity with the two-stream backward"
    CLIPK_BWD_STREAMS=2 timeout 200 python -m pytest tests/test_parity_gpu.py -q -x 2>&1 | tail -5
else
    run "baseline"                      CLIPK_OVERLAP=0
    run "to_f16 under the stats gather" CLIPK_OVERLAP=1
    run "pull-based all-gather"         CLIPK_PEER_GATHER=1
    run "both"                          CLIPK_OVERLAP=1 CLIPK_PEER_GATHER=1
    run "both + two-stream backward"    CLIPK_OVERLAP=1 CLIPK_PEER_GATHER=1 CLIPK_BWD_STREAMS=2
fi
