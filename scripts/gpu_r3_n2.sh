#!/bin/bash
# round 2, third session: the final build on 2 GPUs (default inter-GPU halo, parity inside)
set -u
N=2; TAG=${1:-r03z}; OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 60 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_${TAG}_n2.json 2> $OUT/bench_${TAG}_n2.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_${TAG}_n2.json").read().strip().splitlines()[-1])
    print("N=2 value %.4g ms %.3f parity %s"%(d["value"], d["ms_per_step"], d["parity"]))
    print(d["config"]["inter_gpu_halo"]); print(d["roofline"]["frac"], d["gpu_launches"])
except Exception as e:
    print("bench failed", e); print(open("$OUT/bench_${TAG}_n2.err").read()[-2500:])
PY
