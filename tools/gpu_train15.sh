#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_train_gpu.py -m gpu -q -x 2>&1 | tail -2
MZ_TRAIN_TIMELINE=1 timeout 300 python tools/train_timeline.py 8 128 > $O/t15_timeline.log 2>&1; echo rc=$?
grep -v Warn $O/t15_timeline.log | grep -A12 "^backward" | cut -c1-175
grep "mean\|span" $O/t15_timeline.log
echo "default: $(timeout 120 python tools/train_tower_time.py 8 128 20 2>&1 | tail -1)"
timeout 300 python tools/train_step_target.py 10 1 8 2>&1 | tail -1
echo "no tower kernels: $(MZ_TRAIN_ABLATE=448 timeout 300 python tools/train_step_target.py 10 1 8 2>&1 | tail -1)"
