#!/bin/bash
# round 2, third session: launch list of the final build + the two other configurations
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r03z}
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > $OUT/ncu_launches_$TAG.log 2>&1
tail -1 $OUT/ncu_launches_$TAG.log | cut -c1-200
for c in advection2d sparse3d; do
  timeout 100 python bench.py --config $c --steps 20 --warmup 5 > $OUT/bench_${TAG}_$c.json 2> /dev/null
  python -c "
import json
d=json.loads(open('$OUT/bench_${TAG}_$c.json').read().strip().splitlines()[-1])
print('$c', d['value'], d['ms_per_step'], d['gpu_launches'], d.get('device_busy_fraction'))"
done
