#!/bin/bash
# round 2, first GPU session: new-kernel parity, A/B of the sweep kernels, ncu of the new ones
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=r02a
timeout 600 python -m pytest tests/test_cabi_kernels_gpu.py tests/test_burgers_sim_gpu.py -m gpu -x -q 2>&1 | tail -25 > $OUT/pytest_${TAG}_new.log
tail -6 $OUT/pytest_${TAG}_new.log
for v in v2 v2nopush v1; do
  EXTRA=""; ENVV=""
  [ $v = v2nopush ] && EXTRA="--set pb2/ghost_push=false"
  [ $v = v1 ] && ENVV="PB2_SWEEP_V1=1" && EXTRA="--set pb2/ghost_push=false"
  env $ENVV timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity $EXTRA > $OUT/bench_${TAG}_$v.json 2> $OUT/bench_${TAG}_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_${TAG}_$v.json").read().strip().splitlines()[-1])
    print("$v value %.4g ms %.3f clocks %s"%(d["value"], d["ms_per_step"], d["clocks"]))
    for k,x in d["kernels"].items(): print("   ",k, round(x["ms_total"]/x["launches"],3), round(x.get("gbs",0)))
except Exception as e:
    print("$v failed", e); print(open("$OUT/bench_${TAG}_$v.err").read()[-1500:])
PY
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $OUT/pytest_$TAG.log
tail -4 $OUT/pytest_$TAG.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -c 1500 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sweep_xpair|sweep_chunk' -s 12 -c 3 \
  -o $OUT/prof_$TAG -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > $OUT/ncu_full_$TAG.log 2>&1
tail -2 $OUT/ncu_full_$TAG.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity \
  > $OUT/ncu_launches_$TAG.log 2>&1
tail -2 $OUT/ncu_launches_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
