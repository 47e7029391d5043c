#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python tools/train_tower_check.py 2 16 9 > $O/t7_tower.log 2>&1; echo rc=$?; grep -v Warning $O/t7_tower.log | head -8
echo "default: $(timeout 120 python tools/train_tower_time.py 8 128 20 2>&1 | tail -1)"
echo "no carveout: $(MZ_TRAIN_NO_CARVEOUT=1 timeout 120 python tools/train_tower_time.py 8 128 20 2>&1 | tail -1)"
echo "ablate 15: $(MZ_TRAIN_ABLATE=15 timeout 120 python tools/train_tower_time.py 8 128 20 2>&1 | tail -1)"
for g in 0 1; do timeout 300 python tools/train_step_target.py 10 $g 8 2>&1 | tail -1; done
timeout 600 python -m pytest tests/test_train_gpu.py -m gpu -x -q 2>&1 | tail -5
