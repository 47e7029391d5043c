#!/bin/bash
# MCTS parity tests + a short Gomoku bench line
TAG=${1:-b3}
O=gpurun_out; mkdir -p $O
L=$O/${TAG}_b3.log
{
timeout 600 python -m pytest tests/test_mcts_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 3 --warmup 3 --no-train-step --no-cpu-baseline --no-self-play
} > $L 2>&1
python tools/show_bench.py $L 2>/dev/null || cat $L
