#!/bin/bash
TAG=${1:-bis}
O=gpurun_out; mkdir -p $O tools/bin
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17"
$NV -o tools/bin/umma_bench tools/umma_bench.cu || exit 1
L=$O/${TAG}_bisect.log
{
tools/bin/umma_bench 128 279 128 148 4000 0 0 1 0
tools/bin/umma_bench 128 279 128 148 4000 0 0 1 4
tools/bin/umma_bench 128 279 128 148 4000 0 0 1 8
tools/bin/umma_bench 128 279 128 148 400000 0 0 1 0
for a in 256 319 383 447 511; do
  echo "== ablate $a"
  MZ_CONV_ABLATE=$a MZ_CONV_DEBUG=1 timeout 300 python tools/profile_target.py gomoku 2 2>&1 | grep "conv dbg" | tail -1
done
} > $L 2>&1
cat $L
