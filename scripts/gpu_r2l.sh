#!/bin/bash
# A/B of the FP64 trims on one box, 20 steps each, interleaved twice
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02l}
run() { # name
  local v=$1; shift
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-parity "$@" > $OUT/bench_${TAG}_$v.json 2> $OUT/bench_${TAG}_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_${TAG}_$v.json").read().strip().splitlines()[-1])
    c=d["clocks"]
    print("$v value %.4g ms %.3f sm_mhz %s power %s"%(d["value"], d["ms_per_step"], c["sm_mhz"], c.get("power_w_median")), " ".join("%s %.3f"%(k[6:12],x["ms_total"]/x["launches"]) for k,x in d["kernels"].items()))
except Exception as e:
    print("$v failed", e); print(open("$OUT/bench_${TAG}_$v.err").read()[-800:])
PY
}
build() { rm -f parthenon_b200/csrc/burgers_sweep.o; make -C parthenon_b200/csrc -s -j8 EXTRA="$1" > /dev/null 2>&1 || echo build failed; }
for rep in 1 2; do
  build ""; run both_$rep
  build "-DPB2_INT_LIMITER=0 -DPB2_POW2_SCALE=0"; run none_$rep
  build "-DPB2_INT_LIMITER=0"; run scale_only_$rep
  build "-DPB2_POW2_SCALE=0"; run limiter_only_$rep
done
