#!/bin/bash
# round 2, GPU session F: 64-byte aligned interiors, neighbour prefetch in x; other configs
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02f}
timeout 600 python -m pytest tests/test_cabi_kernels_gpu.py tests/test_burgers_sim_gpu.py -m gpu -x -q 2>&1 | tail -8 > $OUT/pytest_${TAG}_new.log
tail -3 $OUT/pytest_${TAG}_new.log
run() { # name, extra bench args
  local v=$1; shift
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity "$@" > $OUT/bench_${TAG}_$v.json 2> $OUT/bench_${TAG}_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_${TAG}_$v.json").read().strip().splitlines()[-1])
    c=d["clocks"]
    print("$v value %.4g ms %.3f sm_mhz %s power %s %s"%(d["value"], d["ms_per_step"], c["sm_mhz"], c.get("power_w_median"), c["reasons"]))
    for k,x in d["kernels"].items(): print("   ",k, round(x["ms_total"]/x["launches"],3), round(x.get("gbs",0)))
except Exception as e:
    print("$v failed", e); print(open("$OUT/bench_${TAG}_$v.err").read()[-1500:])
PY
}
run lazy
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sweep_xpair|sweep_chunk' -s 12 -c 3 \
  -o $OUT/prof_$TAG -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > $OUT/ncu_full_$TAG.log 2>&1
tail -1 $OUT/ncu_full_$TAG.log
for v in "-DPB2_XPAIR_MINB=3"; do
  rm -f parthenon_b200/csrc/burgers_sweep.o
  make -C parthenon_b200/csrc -s -j8 EXTRA="$v" > /dev/null 2>&1 || { echo build failed $v; continue; }
  run "var$(echo $v | tr -c 'A-Za-z0-9\n' '_')"
done
