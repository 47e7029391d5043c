#!/bin/bash
timeout 600 python -m pytest tests/test_train_gpu.py tests/test_training.py -m gpu -q -x 2>&1 | grep -E "^E|passed|failed" | head -10
timeout 300 python tools/train_step_target.py 10 1 8 2>&1 | tail -1
echo "torch adam: $(MZ_FAST_ADAM=0 timeout 300 python tools/train_step_target.py 10 1 8 2>&1 | tail -1)"
