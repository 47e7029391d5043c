#!/bin/bash
# quick check of a conv-kernel change: role timing under ablations, conv parity tests, short bench
TAG=${1:-q}
O=gpurun_out; mkdir -p $O
L=$O/${TAG}_quick.log
{
for a in ${ABLATES:-383 319 256 768}; do
  echo "== ablate $a"
  MZ_CONV_ABLATE=$a MZ_CONV_DEBUG=1 timeout 300 python tools/profile_target.py gomoku 2 2>&1 | grep "conv dbg" | tail -1
done
timeout 600 python -m pytest tests/test_conv_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 3 --warmup 3 --no-train-step --no-cpu-baseline
} > $L 2>&1
cat $L
