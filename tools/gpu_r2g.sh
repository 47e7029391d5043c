#!/bin/bash
TAG=${1:-r2g}
O=gpurun_out; mkdir -p $O
for w in tictactoe cartpole; do MZ_MLP_DEBUG=1 timeout 120 python tools/profile_target.py $w 6 2>&1 | grep "search dbg" | tail -1; done
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/${TAG}_pytest_gpu.log; cat $O/${TAG}_pytest_gpu.log
for w in tictactoe cartpole; do
 for f in 0 1; do
  if [ $f = 0 ]; then export MZ_NO_FUSED_SEARCH=1; else unset MZ_NO_FUSED_SEARCH; fi
  timeout 300 python bench.py --workload $w --steps 30 --warmup 3 --no-cpu-baseline --no-train-step --no-self-play --no-configs > $O/${TAG}_bench_${w}_f$f.json 2>$O/${TAG}_bench_${w}_f$f.err
  tail -2 $O/${TAG}_bench_${w}_f$f.err; echo "fused $f"; python tools/show_bench.py $O/${TAG}_bench_${w}_f$f.json 2>/dev/null
 done
done
