#!/bin/bash
# full GPU suite on one B200
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02v}
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_$TAG.log 2>&1; tail -15 $OUT/pytest_$TAG.log
