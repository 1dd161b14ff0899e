#!/bin/bash
# one GPU, final state of round 2: the whole -m gpu suite, smoke(), the default bench line, ncu launch lists of the
# same kernels (plain run first), and the panel-budget A/B at the 8-GPU shard shape
set -u
mkdir -p gpurun_out
echo "=== gpu suite"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "=== bench N=1 (default flags)"
timeout 600 python bench.py > gpurun_out/bench_r2z_n1.json 2> gpurun_out/bench_r2z_n1.err || tail -20 gpurun_out/bench_r2z_n1.err
python tests/tools/show_bench.py gpurun_out/bench_r2z_n1.json
echo "=== panel budget at the 8-GPU shard shape (backward of one rank, us)"
for mb in 192 288 96; do echo "CLIPK_PANEL_MB=$mb"; CLIPK_PANEL_MB=$mb SHARD_TIME=1 timeout 100 python tests/tools/shard_step.py 4096 32768 512 1 2>&1 | tail -1; done
echo "=== ncu launch lists"
for shape in "32768 32768 512" "4096 32768 512"; do
  tag=$(echo $shape | tr ' ' 'x')
  python tests/tools/shard_step.py $shape 2 > gpurun_out/plainz_$tag.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_final_$tag.csv \
      python tests/tools/shard_step.py $shape 2 > gpurun_out/ncu_lz_$tag.log 2>&1
  python tests/tools/ncu_summary.py gpurun_out/launches_final_$tag.csv | head -14
done
