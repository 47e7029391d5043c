#!/bin/bash
# compute-sanitizer over one small search per kernel family -> gpurun_out/<tag>_sanitizer.log
TAG=${1:-r2}
O=gpurun_out; mkdir -p $O
L=$O/${TAG}_sanitizer.log
: > $L
for tgt in tree conv_resident conv_dataflow; do
  for tool in memcheck racecheck synccheck; do
    echo "=== compute-sanitizer --tool $tool  python tools/sanitize_target.py $tgt" >> $L
    timeout ${SAN_TIMEOUT:-240} compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py $tgt 2>&1 \
      | grep -v "^$" | tail -40 >> $L
    echo "exit ${PIPESTATUS[0]}" >> $L
  done
done
echo "=== confined tree kernel (mz_pool_set_tree_ctas) under racecheck" >> $L
MZ_SAN_CONFINED=1 timeout ${SAN_TIMEOUT:-240} compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_target.py tree 2>&1 | grep -v "^$" | tail -8 >> $L
cat $L
