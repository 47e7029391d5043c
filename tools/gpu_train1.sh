#!/bin/bash
O=gpurun_out; mkdir -p $O
for sw in 0 1; do
  echo "=== WGRAD_SWAP=$sw fp16 fwd"
  MZ_TRAIN_WGRAD_SWAP=$sw timeout 300 python tools/train_check.py 2 16 9 > $O/t1_check_swap$sw.log 2>&1; echo rc=$?; tail -n 70 $O/t1_check_swap$sw.log | cut -c1-200
done
echo "=== bf16 fwd"
MZ_TRAIN_FWD_BF16=1 timeout 300 python tools/train_check.py 2 16 9 > $O/t1_check_bf16.log 2>&1; echo rc=$?; grep -E "loss|worst|rc=" $O/t1_check_bf16.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-train-step --no-self-play --no-configs 2>$O/t1_bench.err > $O/t1_bench.json; python tools/show_bench.py $O/t1_bench.json
