#!/bin/bash
O=gpurun_out; mkdir -p $O
MZ_TRAIN_TIMELINE=1 timeout 300 python tools/train_timeline.py 8 128 > $O/t11_timeline.log 2>&1; echo rc=$?
grep -v Warn $O/t11_timeline.log | head -75
grep "mean\|span" $O/t11_timeline.log
