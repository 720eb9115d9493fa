#!/bin/bash
# round 2, third session: flux correction of a face field on the device (new tests only)
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r03a}
timeout 200 python -m pytest tests/test_tecomm_gpu.py -m gpu -q -x -k "flux_correction_of_a_face_field" 2>&1 | tail -40 > $OUT/pytest_${TAG}_new.log
tail -30 $OUT/pytest_${TAG}_new.log
