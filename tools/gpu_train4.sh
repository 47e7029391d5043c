#!/bin/bash
O=gpurun_out; mkdir -p $O
TAG=${1:-t4}
timeout 300 python tools/train_tower_check.py 2 16 9 > $O/${TAG}_tower.log 2>&1; echo rc=$?; grep -v Warning $O/${TAG}_tower.log | head -8
timeout 600 python -m pytest tests/test_train_gpu.py -m gpu -x -q 2>&1 | tail -15
for g in 0 1; do timeout 300 python tools/train_step_target.py 10 $g 8 2>&1 | tail -1; done
MZ_NO_PDL=1 timeout 300 python tools/train_step_target.py 10 1 8 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches_train.csv python tools/train_step_target.py 1 0 8 > $O/${TAG}_ncu.log 2>&1
python tools/launch_summary.py $O/${TAG}_launches_train.csv $O/${TAG}_launches_train.txt | head -24
