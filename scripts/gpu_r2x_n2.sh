#!/bin/bash
# 2 GPUs, default bench flags (e2e and parity inside; no CPU leg) and a short strong-scaling run,
# both with the default inter-GPU halo (peer push through the copy engines)
set -u
N=2; TAG=${1:-r02x}; OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_${TAG}_n2.json 2> $OUT/bench_${TAG}_n2.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_${TAG}_n2.json").read().strip().splitlines()[-1])
    print("N=2 value %.4g ms %.3f e2e %.4g (%.1f ms) parity %s"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["parity"]["ok"]))
    print(d["config"]["inter_gpu_halo"]); print(d["roofline"]["frac"], d["gpu_launches"])
except Exception as e:
    print("bench failed", e); print(open("$OUT/bench_${TAG}_n2.err").read()[-3000:])
PY
timeout 300 $TR bench.py --gpus $N --steps 3 --warmup 3 --scaling strong --no-e2e --no-parity --no-cpu-baseline > $OUT/bench_${TAG}_n2_strong.json 2> $OUT/bench_${TAG}_n2_strong.err
python -c "
import json
d=json.loads(open('$OUT/bench_${TAG}_n2_strong.json').read().strip().splitlines()[-1]); print('strong N=2 value %.4g ms %.3f'%(d['value'], d['ms_per_step']), d['config']['inter_gpu_halo'])" || tail -20 $OUT/bench_${TAG}_n2_strong.err
