#!/bin/bash
# experiment: time the stage kernels under different build variants (env PB2_SWEEP_VARIANT)
for v in "$@"; do
  echo "== variant $v"
  PB2_SWEEP_VARIANT=$v timeout 600 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
for k,v in d['kernels'].items(): print('  ',k, round(v['ms_total']/v['launches'],3))
"
done
