#!/bin/bash
# Round-2 GPU visit: GPU suite, smoke, default bench line (all configs), reference arm, sanitizer logs.
# usage: bash tools/gpu_r2.sh <tag> [steps]
TAG=${1:-r2}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $O/${TAG}_pytest_gpu.log 2>&1
echo "smoke exit $?" >> $O/${TAG}_pytest_gpu.log
tail -15 $O/${TAG}_pytest_gpu.log
if [ -z "$SKIP_BENCH" ]; then
timeout 900 python bench.py --steps ${2:-5} --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
echo "bench exit $?" >> $O/${TAG}_bench.err
cat $O/${TAG}_bench.json; tail -5 $O/${TAG}_bench.err
fi
if [ -n "$RUN_REF" ]; then
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
cat $O/${TAG}_bench_reference.json
fi
if [ -n "$RUN_SAN" ]; then bash tools/gpu_sanitize.sh $TAG > /dev/null 2>&1; tail -60 $O/${TAG}_sanitizer.log; fi
