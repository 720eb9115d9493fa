#!/bin/bash
# Runs on the GPU box (under gpurun): GPU parity tests + a short kernel-time table.
# usage: scripts/gpu_quick.sh <tag> [extra bench args]
TAG=${1:-q}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log
tail -4 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d.get("e2e",{}).get("value"), d.get("e2e",{}).get("ms_per_step"))
for k,v in d["kernels"].items(): print("  ",k, round(v["ms_total"]/v["launches"],3), round(v.get("gbs",0)))
print(d["roofline"])
PY
tail -3 gpurun_out/bench_$TAG.err
