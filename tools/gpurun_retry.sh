#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3 / transient): usage tools/gpurun_retry.sh <timeout> '<command>'
T=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out"; exit $rc
done
echo "gave up: pod busy"; exit 3
