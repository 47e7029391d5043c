#!/bin/bash
# training kernels: parity suite, one tower's chain, the K=5 step; A/B of the switches given as arguments (VAR=1 ...)
timeout 600 python -m pytest tests/test_train_gpu.py -m gpu -q -x 2>&1 | tail -3
echo "tower: $(timeout 120 python tools/train_tower_time.py 8 128 20 2>&1 | tail -1 | cut -c50-230)"
timeout 300 python tools/train_step_target.py 10 1 8 2>&1 | tail -1
for v in "$@"; do echo "$v: $(env $v timeout 300 python tools/train_step_target.py 10 1 8 2>&1 | tail -1)"; done
