#!/bin/bash
# Builds the UNMODIFIED reference (parthenon-hpc-lab/parthenon, read-only at /root/reference) with
# its own cmake build (Kokkos OpenMP, no MPI, no HDF5 — recipe of SURVEY.md 8c / BASELINE.md 3)
# outside the repo and installs only the benchmark executable + input deck into baseline/_ref/
# (git-ignored; travels to the GPU box with the snapshot).  bench.py --impl reference runs it.
# Two loop layouts are built: the default SIMDFOR inner loops and MDRANGE (SURVEY.md 8d asks for
# the better of the two).
set -e
REF=${REF:-/root/reference}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/baseline/_ref
JOBS=${JOBS:-8}
[ -d "$REF" ] || { echo "no reference tree at $REF"; exit 1; }
mkdir -p "$OUT"
for layout in simdfor mdrange; do
  B=/tmp/pb2_refbuild_$layout
  if [ -x "$OUT/burgers-benchmark.$layout" ]; then continue; fi
  EXTRA=""
  [ $layout = mdrange ] && EXTRA="-DPAR_LOOP_LAYOUT=MDRANGE_LOOP"
  mkdir -p $B
  (cd $B && cmake -G Ninja "$REF" -DCMAKE_C_COMPILER=/usr/bin/gcc -DCMAKE_CXX_COMPILER=/usr/bin/g++ \
      -DPARTHENON_DISABLE_MPI=ON -DPARTHENON_DISABLE_HDF5=ON -DKokkos_ENABLE_OPENMP=ON \
      -DKokkos_ENABLE_SERIAL=ON -DPARTHENON_ENABLE_PYTHON_MODULE_CHECK=OFF \
      -DCMAKE_BUILD_TYPE=Release -DPARTHENON_LINT_DEFAULT=OFF -DREGRESSION_GOLD_STANDARD_SYNC=OFF \
      $EXTRA > cmake.log 2>&1 && ninja -j $JOBS burgers-benchmark > ninja.log 2>&1)
  cp $B/benchmarks/burgers/burgers-benchmark "$OUT/burgers-benchmark.$layout"
done
cp "$REF/benchmarks/burgers/burgers.pin" "$OUT/burgers.pin"
(cd "$REF" && git rev-parse HEAD 2>/dev/null || echo unknown) > "$OUT/REVISION"
ls -la "$OUT"
