#!/bin/bash
TAG=${1:-r2k}
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_network_gpu.py -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/mlp_batch_latency.py 2>&1 | grep tictactoe
for f in 0 -1; do
  MZ_FUSED_SEARCH=$( [ $f = 0 ] && echo 0 || echo 1 ) timeout 300 python bench.py --workload tictactoe --steps 30 --warmup 3 --no-cpu-baseline --no-train-step --no-self-play --no-configs > $O/${TAG}_bench_ttt_f$f.json 2>$O/${TAG}_bench_ttt_f$f.err
  tail -2 $O/${TAG}_bench_ttt_f$f.err; echo "fused $f"; python tools/show_bench.py $O/${TAG}_bench_ttt_f$f.json 2>/dev/null
done
