#!/bin/bash
# N-GPU weak bench: peer push vs slabs + NCCL, same box, in-run parity
set -u
N=${1:-8}; TAG=${2:-r02q_n$N}; OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
for v in push nccl; do
  EX=""; [ $v = nccl ] && EX="--set pb2/peer_push=false"
  NCCL_DEBUG=WARN timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-cpu-baseline $EX > $OUT/bench_${TAG}_$v.json 2> $OUT/bench_${TAG}_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_${TAG}_$v.json").read().strip().splitlines()[-1])
    print("$v N=$N value %.4g ms %.3f parity %s"%(d["value"], d["ms_per_step"], d["parity"]))
    for k,x in d["kernels"].items(): print("   ",k, x["launches"], round(x["ms_total"]/x["launches"],3), round(x.get("gbs",0)), round(x["share"],3))
    print(d["clocks"])
except Exception as e:
    print("$v bench failed", e); print(open("$OUT/bench_${TAG}_$v.err").read()[-3000:])
PY
done
