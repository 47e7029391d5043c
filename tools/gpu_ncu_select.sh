#!/bin/bash
TAG=${1:-sel}
O=gpurun_out; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:select_kernel -s 380 -c 1 -f -o $O/${TAG}_select_full \
    python tools/profile_target.py gomoku 200 > $O/${TAG}_ncu_select.log 2>&1
tail -3 $O/${TAG}_ncu_select.log
