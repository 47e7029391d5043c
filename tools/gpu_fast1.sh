#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_bench_parity_gpu.py -m gpu -x -q 2>&1 | tail -4
B="--steps 5 --warmup 3 --no-train-step --no-cpu-baseline --no-self-play --no-configs"
timeout 600 python bench.py $B > $O/f1_fast.json 2> $O/f1_fast.err
MZ_CONV_NO_FAST=1 timeout 600 python bench.py $B > $O/f1_nofast.json 2> $O/f1_nofast.err
python - <<'PY'
import json
for n in ('fast', 'nofast'):
    try:
        d = json.loads(open('gpurun_out/f1_%s.json' % n).read().strip().splitlines()[-1])
        r = d['roofline']
        print(n, 'value %.0f' % d['value'], 'ms %.1f' % d['ms_per_step'], 'frac %.3f' % r['frac'], 'launch us %.1f' % r['avg_launch_us'], 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
    except Exception as e:
        print(n, 'failed', e)
PY
MZ_CONV_DEBUG=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-train-step --no-cpu-baseline --no-self-play --no-configs 2>&1 | grep "conv dbg" | tail -2 | cut -c1-400
