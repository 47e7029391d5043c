#!/bin/bash
TAG=${1:-r2h}
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/${TAG}_pytest_gpu.log; cat $O/${TAG}_pytest_gpu.log
for w in tictactoe cartpole; do
 for f in 1 0; do
  if [ $f = 0 ]; then export MZ_NO_PDL=1; else unset MZ_NO_PDL; fi
  timeout 300 python bench.py --workload $w --steps 30 --warmup 3 --no-cpu-baseline --no-train-step --no-self-play --no-configs > $O/${TAG}_bench_${w}_pdl$f.json 2>$O/${TAG}_bench_${w}_pdl$f.err
  tail -2 $O/${TAG}_bench_${w}_pdl$f.err; echo "pdl $f"; python tools/show_bench.py $O/${TAG}_bench_${w}_pdl$f.json 2>/dev/null
 done
done
unset MZ_NO_PDL
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-train-step --no-self-play --no-configs > $O/${TAG}_bench_gomoku.json 2>$O/${TAG}_bench_gomoku.err; tail -2 $O/${TAG}_bench_gomoku.err; python tools/show_bench.py $O/${TAG}_bench_gomoku.json
