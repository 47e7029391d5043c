#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_train_gpu.py -m gpu -q 2>&1 | tail -12 | tee $O/t10_pytest.log
echo "default: $(timeout 120 python tools/train_tower_time.py 8 128 20 2>&1 | tail -1)" | tee $O/t10_time.log
for g in 1 0; do timeout 300 python tools/train_step_target.py 10 $g 8 2>&1 | tail -1 | tee -a $O/t10_time.log; done
MZ_TRAIN_NATIVE=0 timeout 300 python tools/train_step_target.py 10 1 8 2>&1 | tail -1 | tee -a $O/t10_time.log
