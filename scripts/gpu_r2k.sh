#!/bin/bash
# round 2, GPU session K: FP64 trims + deferred unpack (virtual ranks) on one GPU
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02k}
timeout 600 python -m pytest tests/test_cabi_kernels_gpu.py tests/test_burgers_sim_gpu.py tests/test_advection_sim_gpu.py tests/test_sparse_sim_gpu.py -m gpu -x -q 2>&1 | tail -8 > $OUT/pytest_${TAG}_new.log
tail -3 $OUT/pytest_${TAG}_new.log
run() { # name, extra bench args
  local v=$1; shift
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e "$@" > $OUT/bench_${TAG}_$v.json 2> $OUT/bench_${TAG}_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_${TAG}_$v.json").read().strip().splitlines()[-1])
    c=d["clocks"]
    print("$v value %.4g ms %.3f sm_mhz %s power %s %s parity %s"%(d["value"], d["ms_per_step"], c["sm_mhz"], c.get("power_w_median"), c["reasons"], d["parity"]))
    for k,x in d["kernels"].items(): print("   ",k, x["launches"], round(x["ms_total"]/x["launches"],3), round(x.get("gbs",0)))
except Exception as e:
    print("$v failed", e); print(open("$OUT/bench_${TAG}_$v.err").read()[-1500:])
PY
}
run lazy
run vr2 --set pb2/virtual_ranks=2 --no-parity
run vr8 --set pb2/virtual_ranks=8 --no-parity
PB2_TIME_REMESH=1 timeout 300 python scripts/amr_host_profile.py 2>&1 | tail -12
timeout 300 python bench.py --config advection_amr --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('advection_amr value %.4g ms %.3f busy %.2f'%(d['value'], d['ms_per_step'], d['device_busy_fraction']))"
timeout 300 python bench.py --config sparse3d --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('sparse3d value %.4g ms %.3f busy %.2f'%(d['value'], d['ms_per_step'], d['device_busy_fraction']))"
