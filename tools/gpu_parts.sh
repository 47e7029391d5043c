#!/bin/bash
TAG=${1:-parts}
O=gpurun_out; mkdir -p $O
L=$O/${TAG}_parts.log
{
timeout 120 python -m pytest tests/test_conv_gpu.py -m gpu -x -q 2>&1 | tail -8
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "CONV TESTS FAILED"; exit 1; fi
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for pr in ${PARTS:-1 2}; do
  timeout 300 python bench.py --steps 3 --warmup 3 --no-train-step --no-cpu-baseline --no-self-play --parts $pr
done
} > $L 2>&1
cat $L
