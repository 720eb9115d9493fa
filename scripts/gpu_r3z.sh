#!/bin/bash
# round 2, third session, final build on one B200: full GPU suite, smoke, the bench line, the
# adaptive configuration, Parthenon-VIBE as shipped (fast arithmetic only)
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r03z}
timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > $OUT/pytest_${TAG}_full.log
tail -4 $OUT/pytest_${TAG}_full.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --steps 20 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_$TAG.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","parity","cpu_baseline","clocks","gpu_launches")})
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "roofline", d["roofline"]["frac"], d["roofline"]["fp64"]["frac"])
PY
tail -2 $OUT/bench_$TAG.err
timeout 100 python bench.py --config advection_amr --steps 20 --warmup 5 > $OUT/bench_${TAG}_advection_amr.json 2> /dev/null
python -c "
import json
d=json.loads(open('$OUT/bench_${TAG}_advection_amr.json').read().strip().splitlines()[-1])
print('advection_amr', d['value'], d['ms_per_step'], d['config']['blocks'])"
timeout 60 python scripts/pvibe_as_shipped.py fast > $OUT/pvibe_as_shipped_$TAG.json 2> $OUT/pvibe_$TAG.err
python -c "
import json
d=json.loads(open('$OUT/pvibe_as_shipped_$TAG.json').read().strip().splitlines()[-1])
r=d['runs'][0]; print('pvibe', r['zone_cycles_per_wallsecond'], r['wall_s'], r['cycles'], d['vs_published_a100'])"
