#!/bin/bash
# co-residency probes -> profiles (needs tools/bin/libcoresident_probe.so, see tools/coresident_probe.cu)
O=gpurun_out; mkdir -p $O
MZ_TREE_TIMING=1 timeout 300 python tools/coresident_probe.py > $O/coresidency_probe.log 2>&1
tail -25 $O/coresidency_probe.log
