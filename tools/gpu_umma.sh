#!/bin/bash
TAG=${1:-umma}
O=gpurun_out; mkdir -p $O tools/bin
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17"
$NV -o tools/bin/umma_bench tools/umma_bench.cu || exit 1
$NV -o tools/bin/umma_issue_bench tools/umma_issue_bench.cu || exit 1
L=$O/${TAG}_umma.log
{
for data in 0 1; do
 for grid in 1 148; do
  tools/bin/umma_bench 128 279 128 $grid 4000 0 0 $data
  tools/bin/umma_bench 128 279 128 $grid 4000 0 1 $data
  tools/bin/umma_bench 256 279 256 $grid 4000 0 0 $data
  tools/bin/umma_bench 256 279 256 $grid 4000 0 1 $data
 done
done
for w in 1 2 3; do
  tools/bin/umma_bench 128 279 128 148 4000 $w 0 1
  tools/bin/umma_bench 128 279 128 148 4000 $w 1 1
  tools/bin/umma_bench 256 279 256 148 4000 $w 0 1
done
tools/bin/umma_issue_bench 0 148 2000
tools/bin/umma_issue_bench 1 148 2000
} > $L 2>&1
cat $L
