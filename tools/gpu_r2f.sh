#!/bin/bash
TAG=${1:-r2f}
O=gpurun_out; mkdir -p $O
for w in tictactoe cartpole; do MZ_MLP_DEBUG=1 timeout 120 python tools/profile_target.py $w 6 2>&1 | grep "search dbg" | tail -2; done
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/${TAG}_pytest_gpu.log; cat $O/${TAG}_pytest_gpu.log
