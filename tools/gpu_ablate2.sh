#!/bin/bash
# non-instrumented timing of the conv tower launch under ablation switches (ncu duration + cycles per launch)
TAG=${1:-abl2}
O=gpurun_out; mkdir -p $O
L=$O/${TAG}_ablate2.log
for a in ${ABLATES:-0 1 7 15 23 31 63}; do
  MZ_CONV_ABLATE=$a timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max --clock-control none -k regex:conv3x3 -s 2 -c 2 --csv \
     python tools/profile_target.py gomoku 2 2>/dev/null | grep conv3x3 | awk -F'","' -v a=$a '{print "ablate", a, $(NF-2), $NF}' >> $L
done
cat $L
