#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/g1_pytest.log
timeout 900 python bench.py > $O/g1_bench.json 2> $O/g1_bench.err; echo bench rc=$?
python tools/show_bench.py $O/g1_bench.json 2>/dev/null | head -30
