#!/bin/bash
# N-GPU weak bench of the current build, no e2e / cpu legs (short box time), in-run parity kept
set -u
N=${1:-8}; TAG=${2:-r02n_n$N}; OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
NCCL_DEBUG=INFO timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$TAG.json").read().strip().splitlines()[-1])
    print("N=$N value %.4g ms %.3f parity %s"%(d["value"], d["ms_per_step"], d["parity"]))
    for k,x in d["kernels"].items(): print("   ",k, x["launches"], round(x["ms_total"]/x["launches"],3), round(x.get("gbs",0)), round(x["share"],3))
    print(d["roofline"]["frac"], d["clocks"])
except Exception as e:
    print("bench failed", e); print(open("$OUT/bench_$TAG.err").read()[-2500:])
PY
