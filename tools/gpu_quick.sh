#!/bin/bash
# quick check of a conv-kernel change: conv parity tests first (bail out on failure), role timing under
# ablations, the whole GPU suite, a short bench
TAG=${1:-q}
O=gpurun_out; mkdir -p $O
L=$O/${TAG}_quick.log
{
timeout 240 python -m pytest tests/test_conv_gpu.py -m gpu -x -q 2>&1 | tail -25
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "CONV TESTS FAILED"; cat $L; exit 1; fi
for a in ${ABLATES:-0 1 2 4 7 512}; do
  echo "== ablate $a"
  MZ_CONV_ABLATE=$a MZ_CONV_DEBUG=1 timeout 120 python tools/profile_target.py gomoku 2 2>&1 | grep "conv dbg" | tail -1
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 3 --warmup 3 --no-train-step --no-cpu-baseline --no-self-play
} > $L 2>&1
cat $L
