#!/bin/bash
# ncu --set full of the stacked (five calls, persistent CTAs) forward convolution and dgrad of the training step
O=gpurun_out; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tconv_kernel -s 104 -c 2 -f -o $O/s1_tconv_stacked_fwd python tools/train_step_target.py 1 0 8 > $O/s1_ncu_fwd.log 2>&1; tail -1 $O/s1_ncu_fwd.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tconv_kernel -s 122 -c 2 -f -o $O/s1_tconv_stacked_bwd python tools/train_step_target.py 1 0 8 > $O/s1_ncu_bwd.log 2>&1; tail -1 $O/s1_ncu_bwd.log
