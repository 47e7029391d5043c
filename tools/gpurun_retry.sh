#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3 / transient).
# usage: tools/gpurun_retry.sh <timeout> [GPUS=N] '<command>'      (GPUS=2|4|8 asks for that many GPUs of one box)
T=$1; shift
EXTRA=""
if [[ "$1" == GPUS=* ]]; then EXTRA="--gpus ${1#GPUS=}"; shift; fi
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun $EXTRA --timeout $T -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out"; exit $rc
done
echo "gave up: pod busy"; exit 3
