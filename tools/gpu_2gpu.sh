#!/bin/bash
TAG=${1:-r2j}; N=${2:-2}
O=gpurun_out; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > $O/${TAG}_bench_${N}gpu.json 2> $O/${TAG}_bench_${N}gpu.err
tail -5 $O/${TAG}_bench_${N}gpu.err; python tools/show_bench.py $O/${TAG}_bench_${N}gpu.json
python - <<PY
import json
d=json.loads(open('$O/${TAG}_bench_${N}gpu.json').read().strip().splitlines()[-1])
print('strong', d.get('strong_scaling'))
print('train', d.get('train_step'))
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-700
