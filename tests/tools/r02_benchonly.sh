#!/bin/bash
# bash tests/tools/r02_benchonly.sh <N> [pytest -k expression for tests/test_dist_gpu.py]
set -u
N=$1
mkdir -p gpurun_out
if [ -n "${2:-}" ]; then
  echo "=== tests/test_dist_gpu.py -k '$2' with $N GPUs"
  timeout 900 python -m pytest tests/test_dist_gpu.py -q -k "$2" 2>&1 | tail -6
fi
echo "=== bench N=$N"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus "$N" --steps 20 --warmup 5 > gpurun_out/bench_r2i_n$N.json 2> gpurun_out/bench_r2i_n$N.err || tail -30 gpurun_out/bench_r2i_n$N.err
python tests/tools/show_bench.py gpurun_out/bench_r2i_n$N.json
