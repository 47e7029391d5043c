#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_network_gpu.py tests/test_mcts_gpu.py -m gpu -x -q 2>&1 | tail -3
B="--steps 5 --warmup 3 --no-train-step --no-cpu-baseline --no-self-play --no-configs"
for w in tictactoe cartpole; do timeout 600 python bench.py --workload $w $B > $O/m1_$w.json 2> $O/m1_$w.err; python tools/show_bench.py $O/m1_$w.json | head -3 | cut -c1-330; done
