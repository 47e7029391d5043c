#!/bin/bash
# conv kernel role timing under the ablation switches + one full ncu capture of the tower launch
TAG=${1:-abl}
O=gpurun_out
mkdir -p $O
for a in 0 1 2 4 6 7; do
  echo "== ablate $a" >> $O/${TAG}_ablate.log
  MZ_CONV_ABLATE=$a MZ_CONV_DEBUG=1 timeout 300 python tools/profile_target.py gomoku 2 2>&1 | grep "conv dbg" | tail -2 >> $O/${TAG}_ablate.log
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3 -s 2 -c 1 -f -o $O/${TAG}_conv_full \
    python tools/profile_target.py gomoku 2 > $O/${TAG}_ncu_full.log 2>&1
cat $O/${TAG}_ablate.log
tail -4 $O/${TAG}_ncu_full.log
