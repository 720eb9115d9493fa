#!/bin/bash
# multi-GPU session: bench (with in-run parity) + multigpu_check on N GPUs
set -u
N=${1:-2}; TAG=${2:-r02_n$N}; OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
NCCL_DEBUG=INFO timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$TAG.json").read().strip().splitlines()[-1])
    print("N=$N value %.4g ms %.3f e2e %.4g parity %s"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity"]))
    for k,x in d["kernels"].items(): print("   ",k, x["launches"], round(x["ms_total"]/x["launches"],3), round(x.get("gbs",0)), round(x["share"],3))
    print(d["roofline"])
except Exception as e:
    print("bench failed", e); print(open("$OUT/bench_$TAG.err").read()[-2500:])
PY
grep -c "NCCL INFO" $OUT/bench_$TAG.err; grep "NCCL INFO.*nranks\|Init COMPLETE" $OUT/bench_$TAG.err | head -4
timeout 900 $TR scripts/multigpu_check.py > $OUT/multigpu_check_$TAG.txt 2>&1
tail -12 $OUT/multigpu_check_$TAG.txt
if [ "${3:-}" = strong ]; then
  timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --scaling strong --no-e2e --no-parity > $OUT/bench_${TAG}_strong.json 2> $OUT/bench_${TAG}_strong.err
  python -c "
import json
d=json.loads(open('$OUT/bench_${TAG}_strong.json').read().strip().splitlines()[-1]); print('strong N=$N value %.4g ms %.3f'%(d['value'], d['ms_per_step']))"
fi
