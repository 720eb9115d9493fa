#!/bin/bash
# round 2, third session: host path of a remesh after the rebuild work (table pool, parallel
# copy regions, shared plans): phases, configs[2] bench line, the adaptive tests
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r03d}
PB2_TIME_HOST=1 timeout 120 python scripts/amr_host_profile.py > $OUT/amr_host_profile_$TAG.txt 2>&1
grep -E "rebuild base|rebuild 1|remesh:|blocks " $OUT/amr_host_profile_$TAG.txt | tail -8 | cut -c1-330
timeout 200 python bench.py --config advection_amr --steps 20 --warmup 5 > $OUT/bench_${TAG}_advection_amr.json 2> $OUT/bench_${TAG}_advection_amr.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_${TAG}_advection_amr.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","device_busy_fraction")}, d["config"]["blocks"])
PY
tail -2 $OUT/bench_${TAG}_advection_amr.err
timeout 300 python -m pytest tests -m gpu -q -x -k "adaptive or remesh or sparse" 2>&1 | tail -5
