#!/bin/bash
# check of a tree-kernel change: MCTS parity tests first, whole GPU suite, short bench lines of three workloads
TAG=${1:-s}
O=gpurun_out; mkdir -p $O
L=$O/${TAG}_select.log
{
timeout 600 python -m pytest tests/test_mcts_gpu.py -m gpu -x -q 2>&1 | tail -25
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 3 --warmup 3 --no-train-step --no-cpu-baseline --no-self-play
timeout 300 python bench.py --workload tictactoe --steps 10 --warmup 3 --no-train-step --no-cpu-baseline
timeout 300 python bench.py --workload cartpole --steps 10 --warmup 3 --no-train-step --no-cpu-baseline
} > $L 2>&1
cat $L
