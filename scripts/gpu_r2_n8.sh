#!/bin/bash
# N-GPU bench: weak (with in-run parity) + strong
set -u
N=${1:-8}; TAG=${2:-r02_n$N}; OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
NCCL_DEBUG=INFO timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$TAG.json").read().strip().splitlines()[-1])
    print("N=$N value %.4g ms %.3f e2e %.4g parity %s"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity"]))
    for k,x in d["kernels"].items(): print("   ",k, x["launches"], round(x["ms_total"]/x["launches"],3), round(x.get("gbs",0)), round(x["share"],3))
    print(d["roofline"]["frac"], d["clocks"])
except Exception as e:
    print("bench failed", e); print(open("$OUT/bench_$TAG.err").read()[-2500:])
PY
grep "NCCL INFO.*Init COMPLETE" $OUT/bench_$TAG.err | head -2
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --scaling strong --no-e2e --no-parity > $OUT/bench_${TAG}_strong.json 2> $OUT/bench_${TAG}_strong.err
python -c "
import json
d=json.loads(open('$OUT/bench_${TAG}_strong.json').read().strip().splitlines()[-1]); print('strong N=$N value %.4g ms %.3f'%(d['value'], d['ms_per_step']))" || tail -20 $OUT/bench_${TAG}_strong.err
