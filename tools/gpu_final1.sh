#!/bin/bash
# round-2 closing run on one GPU: full GPU suite, training-kernel evidence (timeline, launch list, ncu --set full of the
# stacked forward conv / dgrad / wgrad), micro-benchmarks, default bench
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/h1_pytest.log
MZ_TRAIN_TIMELINE=1 timeout 300 python tools/train_timeline.py 8 128 > $O/h1_timeline.log 2>&1; echo timeline rc=$?
timeout 600 ncu --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/h1_launches_train.csv python tools/train_step_target.py 1 0 8 > $O/h1_ncu.log 2>&1
python tools/launch_summary.py $O/h1_launches_train.csv $O/h1_launches_train.txt | head -8
# one launch each of the stacked (5 calls) prediction chain: skip the warm-up steps' launches
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tconv_kernel -s 1200 -c 2 -f -o $O/h1_tconv_full python tools/train_step_target.py 1 0 8 > $O/h1_ncu_tconv.log 2>&1; tail -2 $O/h1_ncu_tconv.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:twgrad_kernel -s 300 -c 2 -f -o $O/h1_twgrad_full python tools/train_step_target.py 1 0 8 > $O/h1_ncu_twgrad.log 2>&1; tail -2 $O/h1_ncu_twgrad.log
timeout 60 stdbuf -oL tools/bin/issue_bench2 100 > $O/h1_issue_bench2.log 2>&1
timeout 60 tools/bin/ring_bench 100 > $O/h1_ring_bench.log 2>&1
timeout 60 tools/bin/chain_probe > $O/h1_chain_probe.log 2>&1
timeout 900 python bench.py > $O/h1_bench.json 2> $O/h1_bench.err; echo bench rc=$?
python tools/show_bench.py $O/h1_bench.json 2>/dev/null | head -30
