#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python tools/train_tower_check.py 2 16 9 > $O/t8_tower.log 2>&1; echo rc=$?; grep -v Warning $O/t8_tower.log | head -8
timeout 600 python -m pytest tests/test_train_gpu.py -m gpu -q 2>&1 | tail -12
echo "default: $(timeout 120 python tools/train_tower_time.py 8 128 20 2>&1 | tail -1)"
for g in 1; do timeout 300 python tools/train_step_target.py 10 $g 8 2>&1 | tail -1; done
