#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/k1_pytest.log
MZ_TRAIN_TIMELINE=1 timeout 300 python tools/train_timeline.py 8 128 > $O/k1_timeline.log 2>&1; echo timeline rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tconv_kernel -s 240 -c 3 -f -o $O/k1_tconv_full python tools/train_step_target.py 1 0 8 > $O/k1_ncu_tconv.log 2>&1; tail -1 $O/k1_ncu_tconv.log
timeout 600 ncu --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/k1_launches_train.csv python tools/train_step_target.py 1 0 8 > $O/k1_ncu.log 2>&1
python tools/launch_summary.py $O/k1_launches_train.csv $O/k1_launches_train.txt | head -12
timeout 900 python bench.py > $O/k1_bench.json 2> $O/k1_bench.err; echo bench rc=$?
python tools/show_bench.py $O/k1_bench.json 2>/dev/null | head -30
timeout 900 python bench.py --impl reference > $O/k1_bench_reference.json 2> $O/k1_bench_reference.err; echo ref rc=$?; cut -c1-400 $O/k1_bench_reference.json
