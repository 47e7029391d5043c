#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_train_gpu.py -m gpu -q -x 2>&1 | tail -8
timeout 300 python tools/train_step_target.py 10 1 8 2>&1 | tail -1
echo "no stacking: $(MZ_TRAIN_NO_STACK=1 timeout 300 python tools/train_step_target.py 10 1 8 2>&1 | tail -1)"
