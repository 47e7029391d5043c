#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 ncu --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/t20_launches_train.csv python tools/train_step_target.py 1 0 8 > $O/t20_ncu.log 2>&1
python tools/launch_summary.py $O/t20_launches_train.csv $O/t20_launches_train.txt | head -60
