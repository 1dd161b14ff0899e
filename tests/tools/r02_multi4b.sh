#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== remaining tests of tests/test_dist_gpu.py with 4 GPUs (2- and 4-rank cases)"
timeout 1500 python -m pytest tests/test_dist_gpu.py -q -k "peer_memory or ddp" 2>&1 | tail -25
echo "=== bench N=4"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_r2_n4.json 2> gpurun_out/bench_r2_n4.err || tail -30 gpurun_out/bench_r2_n4.err
python tests/tools/show_bench.py gpurun_out/bench_r2_n4.json
