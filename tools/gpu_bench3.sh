#!/bin/bash
# full GPU suite + short benches of three workloads
TAG=${1:-b3}
O=gpurun_out; mkdir -p $O
L=$O/${TAG}_bench3.log
{
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 3 --warmup 3 --no-train-step --no-cpu-baseline --no-self-play
timeout 600 python bench.py --workload atari --steps 3 --warmup 3 --no-train-step --no-cpu-baseline --no-self-play
timeout 600 python bench.py --workload tictactoe --steps 5 --warmup 3 --no-train-step --no-cpu-baseline --no-self-play
} > $L 2>&1
cat $L
