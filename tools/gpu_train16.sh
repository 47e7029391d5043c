#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_train_gpu.py -m gpu -q -x 2>&1 | tail -4
timeout 300 python tools/train_step_target.py 10 1 8 2>&1 | tail -1
echo "loop heads: $(MZ_TRAIN_BATCHED_HEADS=0 timeout 300 python tools/train_step_target.py 10 1 8 2>&1 | tail -1)"
echo "no tower kernels: $(MZ_TRAIN_ABLATE=448 timeout 300 python tools/train_step_target.py 10 1 8 2>&1 | tail -1)"
