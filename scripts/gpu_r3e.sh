#!/bin/bash
# round 2, third session: noise level and OpenMP sensitivity of the adaptive bench line
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r03e}
run () { # name, env...
  local name=$1; shift
  env "$@" PB2_TIME_REMESH=1 timeout 100 python bench.py --config advection_amr --steps 20 --warmup 5 > $OUT/bench_${TAG}_$name.json 2> $OUT/bench_${TAG}_$name.err
  python - <<PY
import json,re
d=json.loads(open("$OUT/bench_${TAG}_$name.json").read().strip().splitlines()[-1])
t=[l for l in open("$OUT/bench_${TAG}_$name.err") if l.startswith("rebuild")]
def avg(key):
    v=[float(re.search(key+r" ([0-9.]+)",l).group(1)) for l in t[-20:] if re.search(key+r" ([0-9.]+)",l)]
    return round(sum(v)/max(len(v),1),2)
print("$name", round(d["ms_per_step"],2), "ms/step; last rebuilds: tables", avg("tables"), "copy regions", avg("copy regions"), "copy table", avg("copy table"), "prores", avg("prores"), "plan", avg("plan"))
PY
}
python -c "import os; print('affinity', len(os.sched_getaffinity(0)), 'cpu.max', open('/sys/fs/cgroup/cpu.max').read().strip() if os.path.exists('/sys/fs/cgroup/cpu.max') else None)"
run A X=1
run serial_prores PB2_SERIAL_PRORES=1
run omp4 OMP_NUM_THREADS=4
run passive OMP_WAIT_POLICY=passive
run A2 X=1
