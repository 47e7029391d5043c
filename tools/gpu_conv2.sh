#!/bin/bash
TAG=${1:-c2}
O=gpurun_out; mkdir -p $O
L=$O/${TAG}_conv2.log
{
timeout 240 python -m pytest tests/test_conv_gpu.py -m gpu -x -q 2>&1 | tail -6
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "CONV TESTS FAILED"; exit 1; fi
MZ_CONV_DEBUG=1 timeout 120 python tools/profile_target.py atari 2 2>&1 | grep "conv dbg" | tail -1
MZ_CONV_DEBUG=1 timeout 120 python tools/profile_target.py gomoku 2 1024 2>&1 | grep "conv dbg" | tail -1
timeout 300 python bench.py --workload atari --steps 5 --warmup 3 --no-train-step --no-cpu-baseline
timeout 300 python bench.py --steps 3 --warmup 3 --no-train-step --no-cpu-baseline --no-self-play
} > $L 2>&1
cat $L
