#!/bin/bash
# headline sweep: SMs of a tower launch x persistent CTAs of the confined tree kernel (same box, back to back)
O=gpurun_out; mkdir -p $O
B="--steps 5 --warmup 3 --no-train-step --no-cpu-baseline --no-self-play --no-configs --parts 2"
for lim in 140 144 146; do for tc in 2 4 6; do
  MZ_TREE_CTAS=$tc timeout 300 python bench.py $B --cta-limit $lim > $O/sw_${lim}_${tc}.json 2>/dev/null
  python - <<PY
import json
try:
    d=json.loads(open('$O/sw_${lim}_${tc}.json').read().strip().splitlines()[-1])
    print('cta_limit $lim tree_ctas $tc: value %.0f  ms %.1f  conv launch %.1f us  clk %s' % (d['value'], d['ms_per_step'], d['roofline']['avg_launch_us'], d['clocks']['sm_mhz']))
except Exception as e:
    print('cta_limit $lim tree_ctas $tc: failed', e)
PY
done; done
