#!/bin/bash
# round 2, GPU session I: full suite with the device-progress overlap (virtual ranks), fp64 instruction counts
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02i}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $OUT/pytest_$TAG.log
tail -5 $OUT/pytest_$TAG.log
timeout 300 ncu --metrics smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'sweep_xpair|sweep_chunk' -s 12 -c 6 --csv \
  --log-file $OUT/fp64_inst_$TAG.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > /dev/null 2>&1
tail -8 $OUT/fp64_inst_$TAG.csv | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --set pb2/virtual_ranks=2 > $OUT/bench_${TAG}_vr2.json 2> $OUT/bench_${TAG}_vr2.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_${TAG}_vr2.json").read().strip().splitlines()[-1])
    print("vr2 value %.4g ms %.3f"%(d["value"], d["ms_per_step"]))
    for k,x in d["kernels"].items(): print("   ",k, x["launches"], round(x["ms_total"]/x["launches"],3), round(x.get("gbs",0)))
except Exception as e:
    print("vr2 failed", e); print(open("$OUT/bench_${TAG}_vr2.err").read()[-1500:])
PY
