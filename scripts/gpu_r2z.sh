#!/bin/bash
# round 2, last GPU session on one B200: the tests added last, the full bench line, smoke, launch list
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02z}
timeout 600 python -m pytest tests -m gpu -q -k "many_partitions or peer_push_primitives" 2>&1 | tail -15 > $OUT/pytest_${TAG}_new.log
tail -4 $OUT/pytest_${TAG}_new.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_$TAG.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","parity","cpu_baseline","clocks","gpu_launches")})
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "roofline", d["roofline"], d["cycle_roofline"])
PY
tail -3 $OUT/bench_$TAG.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > $OUT/ncu_launches_$TAG.log 2>&1
tail -1 $OUT/ncu_launches_$TAG.log | cut -c1-300
