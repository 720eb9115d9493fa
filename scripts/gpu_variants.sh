#!/bin/bash
# experiment: rebuild the sweep kernels with extra nvcc flags on the GPU box and time the bench.
# usage: scripts/gpu_variants.sh "<flags A>" "<flags B>" ...   ("" = the default build)
for v in "$@"; do
  echo "== variant [$v]"
  rm -f parthenon_b200/csrc/burgers_sweep.o parthenon_b200/csrc/exchange.o
  make -C parthenon_b200/csrc -s -j8 EXTRA="$v" > /dev/null 2>&1 || { echo build failed; continue; }
  timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
for k,v in d['kernels'].items(): print('  ',k, round(v['ms_total']/v['launches'],3))
"
done
