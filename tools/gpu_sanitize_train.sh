#!/bin/bash
# compute-sanitizer over one small training step on the kernels of csrc/train.cu -> gpurun_out/<tag>_sanitizer_train.log
TAG=${1:-r2}
O=gpurun_out; mkdir -p $O
L=$O/${TAG}_sanitizer_train.log
: > $L
for tool in memcheck synccheck racecheck; do
  echo "=== compute-sanitizer --tool $tool  python tools/sanitize_target.py train" >> $L
  timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool $tool --print-limit 12 python tools/sanitize_target.py train 2>&1 \
    | grep -v "^$" | grep -v Warning | tail -30 >> $L
  echo "exit ${PIPESTATUS[0]}" >> $L
done
cat $L | cut -c1-220
