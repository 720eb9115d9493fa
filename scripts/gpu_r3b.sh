#!/bin/bash
# round 2, third session: where the host time of an adaptive cycle goes (configs[2])
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r03b}
PB2_TIME_HOST=1 timeout 120 python scripts/amr_host_profile.py > $OUT/amr_host_profile_$TAG.txt 2>&1
tail -60 $OUT/amr_host_profile_$TAG.txt | cut -c1-220
timeout 120 python -m pytest tests/test_advection_sim_gpu.py -m gpu -q -x -k "adaptive" 2>&1 | tail -5
