#!/bin/bash
TAG=${1:-r2i}
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $O/${TAG}_pytest_gpu.log; cat $O/${TAG}_pytest_gpu.log
timeout 300 python tools/mlp_batch_latency.py > $O/${TAG}_mlp_batch_latency.log 2>&1; cat $O/${TAG}_mlp_batch_latency.log | tail -12
MZ_CONV_DEBUG=1 timeout 120 python tools/profile_target.py gomoku 2 1024 2>&1 | grep "conv dbg" | tail -3 > $O/${TAG}_conv_role_timing.log; cat $O/${TAG}_conv_role_timing.log
timeout 300 python tools/dropin_latency.py > $O/${TAG}_dropin_latency.log 2>&1; cat $O/${TAG}_dropin_latency.log | tail -5
MZ_FUSED_SEARCH=1 timeout 300 python tools/dropin_latency.py cartpole tictactoe > $O/${TAG}_dropin_latency_fused.log 2>&1; cat $O/${TAG}_dropin_latency_fused.log | tail -3
timeout 900 python tools/fp16_search_stats.py $O/${TAG}_fp16_search_stats.json 512 256 32 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
tail -3 $O/${TAG}_bench.err; python tools/show_bench.py $O/${TAG}_bench.json
