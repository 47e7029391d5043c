#!/bin/bash
TAG=${1:-r2d}
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_network_gpu.py -m gpu -x -q 2>&1 | tail -25
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/${TAG}_pytest_gpu.log; cat $O/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for w in tictactoe cartpole; do
timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-train-step --no-self-play --no-configs > $O/${TAG}_bench_$w.json 2>$O/${TAG}_bench_$w.err
tail -3 $O/${TAG}_bench_$w.err; python tools/show_bench.py $O/${TAG}_bench_$w.json
done
