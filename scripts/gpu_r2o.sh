#!/bin/bash
# peer push: single-GPU tests of the path (virtual ranks) on GPU 0, then the 2-GPU bench with and
# without it (in-run parity: N-rank result bit-identical to one rank) and the multi-GPU parity script
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02o}; N=${2:-2}
timeout 600 python -m pytest tests -m gpu -x -q -k "multiblock_strict or lazy_local_ghosts_match or overlapped_halo or logical_coordinate or non_cell_centred_exchange_bit_exact or multilevel_exchange_matches or forest" > $OUT/pytest_$TAG.log 2>&1; tail -5 $OUT/pytest_$TAG.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
for v in ce nccl sm; do
  EX=""; [ $v = nccl ] && EX="--set pb2/peer_push=false"; [ $v = sm ] && EX="--set pb2/peer_push_mode=sm"
  timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-cpu-baseline $EX > $OUT/bench_${TAG}_$v.json 2> $OUT/bench_${TAG}_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_${TAG}_$v.json").read().strip().splitlines()[-1])
    print("$v N=$N value %.4g ms %.3f parity %s"%(d["value"], d["ms_per_step"], d["parity"]))
    for k,x in d["kernels"].items(): print("   ",k, x["launches"], round(x["ms_total"]/x["launches"],3), round(x.get("gbs",0)), round(x["share"],3))
except Exception as e:
    print("$v bench failed", e); print(open("$OUT/bench_${TAG}_$v.err").read()[-3000:])
PY
  grep -h "pb2\]" $OUT/bench_${TAG}_$v.err | head -3
done
timeout 600 $TR scripts/multigpu_check.py > $OUT/multigpu_check_$TAG.txt 2>&1; tail -4 $OUT/multigpu_check_$TAG.txt
