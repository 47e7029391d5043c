#!/bin/bash
# MLP configs: launch chain with 1 / 2 / 4 sub-batches in flight (tree kernels of one overlap the MLP kernel of another)
TAG=${1:-r2e}
O=gpurun_out; mkdir -p $O
for w in tictactoe cartpole; do
 for parts in 1 2 4; do
  MZ_NO_FUSED_SEARCH=1 timeout 300 python bench.py --workload $w --parts $parts --steps 30 --warmup 3 --no-cpu-baseline --no-train-step --no-self-play --no-configs > $O/${TAG}_bench_${w}_p$parts.json 2>$O/${TAG}_bench_${w}_p$parts.err
  tail -2 $O/${TAG}_bench_${w}_p$parts.err; echo "parts $parts"; python tools/show_bench.py $O/${TAG}_bench_${w}_p$parts.json | head -1
 done
done
