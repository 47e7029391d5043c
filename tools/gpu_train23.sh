#!/bin/bash
timeout 600 python -m pytest tests/test_train_gpu.py -m gpu -q -x 2>&1 | tail -3
echo "tower: $(timeout 120 python tools/train_tower_time.py 8 128 20 2>&1 | tail -1 | cut -c50-230)"
timeout 300 python tools/train_step_target.py 10 1 8 2>&1 | tail -1
echo "no persist: $(MZ_TRAIN_NO_PERSIST=1 timeout 300 python tools/train_step_target.py 10 1 8 2>&1 | tail -1)"
