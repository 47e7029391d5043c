#!/bin/bash
O=gpurun_out; mkdir -p $O
for ab in 0 1 2 4 8 15; do echo "ablate $ab: $(MZ_TRAIN_ABLATE=$ab timeout 120 python tools/train_tower_time.py 8 128 20 2>&1 | tail -1)"; done
echo "no pdl: $(MZ_NO_PDL=1 timeout 120 python tools/train_tower_time.py 8 128 20 2>&1 | tail -1)"
timeout 600 ncu --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/t6_launches_train_warm.csv python tools/train_step_target.py 1 0 8 > $O/t6_ncu.log 2>&1
python tools/launch_summary.py $O/t6_launches_train_warm.csv $O/t6_launches_train_warm.txt | head -12
