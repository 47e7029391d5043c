#!/bin/bash
O=gpurun_out; mkdir -p $O
for ab in 8 1 9; do
MZ_TRAIN_ABLATE=$ab MZ_TRAIN_TIMELINE=1 timeout 300 python tools/train_timeline.py 8 128 > $O/t12_timeline_ab$ab.log 2>&1; echo "ablate $ab rc=$?"
grep -v Warn $O/t12_timeline_ab$ab.log | sed -n 9,14p | cut -c1-150
grep "mean\|span" $O/t12_timeline_ab$ab.log | head -3
done
