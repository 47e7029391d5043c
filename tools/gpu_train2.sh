#!/bin/bash
O=gpurun_out; mkdir -p $O
for sw in 0 1; do
  echo "=== WGRAD_SWAP=$sw"
  MZ_TRAIN_WGRAD_SWAP=$sw timeout 300 python tools/train_tower_check.py 2 16 9 > $O/t2_tower_swap$sw.log 2>&1; echo rc=$?; tail -n 60 $O/t2_tower_swap$sw.log | cut -c1-160
done
