#!/bin/bash
# multi-GPU check of round 2: bash tests/tools/r02_multi.sh <N> [bench-only]
# the NCCL / peer-memory parity tests of tests/test_dist_gpu.py that fit N GPUs, then bench.py at 1 (first call only) and N
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
nvidia-smi topo -m 2>/dev/null | head -12
if [ "${2:-}" != "bench-only" ]; then
  echo "=== tests/test_dist_gpu.py with $N GPUs"
  timeout 1500 python -m pytest tests/test_dist_gpu.py -q -x 2>&1 | tail -12
fi
if [ "$N" = 2 ]; then
  echo "=== bench N=1"
  timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err || tail -20 gpurun_out/bench_r2_n1.err
  python tests/tools/show_bench.py gpurun_out/bench_r2_n1.json
fi
echo "=== bench N=$N"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus "$N" --steps 20 --warmup 5 > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err || tail -30 gpurun_out/bench_r2_n$N.err
python tests/tools/show_bench.py gpurun_out/bench_r2_n$N.json
