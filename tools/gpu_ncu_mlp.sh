#!/bin/bash
TAG=${1:-mlp}
O=gpurun_out; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_recurrent_tc -s 30 -c 1 -f -o $O/${TAG}_mlp_tc_full \
    python tools/profile_target.py tictactoe 25 > $O/${TAG}_ncu_mlp.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_recurrent_tc -s 60 -c 1 -f -o $O/${TAG}_mlp_tc_cartpole_full \
    python tools/profile_target.py cartpole 50 > $O/${TAG}_ncu_mlp2.log 2>&1
tail -2 $O/${TAG}_ncu_mlp.log $O/${TAG}_ncu_mlp2.log
