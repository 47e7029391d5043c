#!/bin/bash
TAG=${1:-atari}
O=gpurun_out; mkdir -p $O
L=$O/${TAG}_atari.log
{
timeout 300 python -m pytest tests/test_conv_gpu.py -m gpu -x -q 2>&1 | tail -25
timeout 300 python bench.py --workload atari --steps 5 --warmup 3 --no-train-step --no-cpu-baseline
} > $L 2>&1
cat $L
