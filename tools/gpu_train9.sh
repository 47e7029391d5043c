#!/bin/bash
for ab in 64 128 256 320 15; do echo "ablate $ab: $(MZ_TRAIN_ABLATE=$ab timeout 120 python tools/train_tower_time.py 8 128 20 2>&1 | tail -1 | cut -c50-160)"; done
echo "nopdl 64: $(MZ_NO_PDL=1 MZ_TRAIN_ABLATE=64 timeout 120 python tools/train_tower_time.py 8 128 20 2>&1 | tail -1 | cut -c50-160)"
echo "nopdl 128: $(MZ_NO_PDL=1 MZ_TRAIN_ABLATE=128 timeout 120 python tools/train_tower_time.py 8 128 20 2>&1 | tail -1 | cut -c50-160)"
