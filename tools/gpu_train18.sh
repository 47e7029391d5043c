#!/bin/bash
O=gpurun_out; mkdir -p $O
MZ_TRAIN_TIMELINE=1 timeout 300 python tools/train_timeline.py 8 128 > $O/t18_timeline.log 2>&1; echo rc=$?
grep -v Warn $O/t18_timeline.log | grep "wgrad" | head -8 | cut -c1-200
