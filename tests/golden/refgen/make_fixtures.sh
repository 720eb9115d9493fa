#!/bin/bash
# Fixture generator (test infrastructure).  Runs the UNMODIFIED reference to produce the
# golden files committed under tests/golden/.  Needs:
#   REF        reference source tree (default /root/reference, read-only)
#   REF_BUILD  an out-of-tree CPU build of the reference (libparthenon.a + Kokkos libs),
#              made with the recipe in SURVEY.md §8c (cmake, Kokkos OpenMP+Serial, no MPI,
#              no HDF5, -O3, g++ 13.3).  Default /tmp/survey_ref_build_own.
# Nothing is written into REF; all scratch goes to $WORK (default /tmp/pb2_refgen).
set -euo pipefail
REF=${REF:-/root/reference}
REF_BUILD=${REF_BUILD:-/tmp/survey_ref_build_own}
WORK=${WORK:-/tmp/pb2_refgen}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=${OUT:-$(cd "$HERE/.." && pwd)}
mkdir -p "$WORK"

INC="-I$REF/src -I$REF_BUILD/src/generated -I$REF_BUILD/Kokkos -I$REF_BUILD/Kokkos/core/src \
 -I$REF/external/Kokkos/core/src -I$REF/external/Kokkos/tpls/desul/include \
 -I$REF_BUILD/Kokkos/containers/src -I$REF/external/Kokkos/containers/src \
 -I$REF_BUILD/Kokkos/algorithms/src -I$REF/external/Kokkos/algorithms/src \
 -I$REF_BUILD/Kokkos/simd/src -I$REF/external/Kokkos/simd/src"
LIBS="$REF_BUILD/src/libparthenon.a $REF_BUILD/Kokkos/containers/src/libkokkoscontainers.a \
 $REF_BUILD/Kokkos/core/src/libkokkoscore.a $REF_BUILD/Kokkos/simd/src/libkokkossimd.a -ldl -fopenmp -lpthread"
CXX=/usr/bin/g++
FLAGS="-O3 -DNDEBUG -std=c++17 -fopenmp -DKOKKOS_DEPENDENCE"

if [ ! -x "$WORK/burgers_dump" ] || [ "$HERE/burgers_dump_main.cpp" -nt "$WORK/burgers_dump" ]; then
  B=$REF/benchmarks/burgers
  $CXX $FLAGS $INC -I$B "$HERE/burgers_dump_main.cpp" $B/burgers_driver.cpp \
     $B/burgers_package.cpp $B/parthenon_app_inputs.cpp $LIBS -o "$WORK/burgers_dump"
fi

if [ ! -x "$WORK/advection_dump" ] || [ "$HERE/advection_dump_main.cpp" -nt "$WORK/advection_dump" ]; then
  A=$REF/example/advection
  $CXX $FLAGS $INC -I$A "$HERE/advection_dump_main.cpp" $A/advection_driver.cpp \
     $A/advection_package.cpp $A/parthenon_app_inputs.cpp $LIBS -o "$WORK/advection_dump"
fi

if [ ! -x "$WORK/sparse_dump" ] || [ "$HERE/sparse_dump_main.cpp" -nt "$WORK/sparse_dump" ]; then
  S=$REF/example/sparse_advection
  $CXX $FLAGS $INC -I$S "$HERE/sparse_dump_main.cpp" $S/sparse_advection_driver.cpp \
     $S/sparse_advection_package.cpp $S/parthenon_app_inputs.cpp $LIBS -o "$WORK/sparse_dump"
fi

if [ ! -x "$WORK/tecomm_dump" ] || [ "$HERE/tecomm_dump_main.cpp" -nt "$WORK/tecomm_dump" ]; then
  $CXX $FLAGS $INC "$HERE/tecomm_dump_main.cpp" $LIBS -o "$WORK/tecomm_dump"
fi

if [ ! -x "$WORK/forest_dump" ] || [ "$HERE/forest_dump_main.cpp" -nt "$WORK/forest_dump" ]; then
  $CXX $FLAGS $INC "$HERE/forest_dump_main.cpp" $LIBS -o "$WORK/forest_dump"
fi

export OMP_NUM_THREADS=${OMP_NUM_THREADS:-8} OMP_PROC_BIND=false

run_burgers () { # name nx nb nscal recon nlim extra...
  local name=$1 nx=$2 nb=$3 nscal=$4 recon=$5 nlim=$6; shift 6
  local d="$WORK/$name"; rm -rf "$d"; mkdir -p "$d"; cd "$d"
  PB2_DUMP_PREFIX="$d/U" "$WORK/burgers_dump" -i "$REF/benchmarks/burgers/burgers.pin" \
    parthenon/mesh/nx1=$nx parthenon/mesh/nx2=$nx parthenon/mesh/nx3=$nx \
    parthenon/meshblock/nx1=$nb parthenon/meshblock/nx2=$nb parthenon/meshblock/nx3=$nb \
    parthenon/mesh/refinement=none parthenon/mesh/numlevel=1 \
    parthenon/time/nlim=$nlim parthenon/time/tlim=1e9 parthenon/output0/dt=-1 \
    parthenon/output1/dt=1e-9 burgers/num_scalars=$nscal burgers/recon=$recon "$@" \
    > run.log 2>&1
  python3 "$HERE/pack_dumps.py" "$d" "$OUT/$name.npz"
  cp "$d/burgers.out1.hst" "$OUT/$name.hst"
}

# small uniform cases: full fields (ghosts included) after cycles 0..nlim
run_burgers burgers_u16_b8_s1_weno5   16  8 1 weno5  3
run_burgers burgers_u16_b8_s1_linear  16  8 1 linear 3 parthenon/mesh/nghost=2
# physical boundaries: outflow in x1, reflecting in x2, periodic in x3
if [ -z "${SKIP_BC:-}" ]; then
run_burgers burgers_u16_b8_s1_weno5_bc 16 8 1 weno5 3 parthenon/mesh/ix1_bc=outflow \
  parthenon/mesh/ox1_bc=outflow parthenon/mesh/ix2_bc=reflecting parthenon/mesh/ox2_bc=reflecting
fi
# static-refinement (multilevel) cases: restrict / prolongate ghost fill (cycle 0) and, after
# the first cycles, flux correction.  The deck is the reference deck plus
# <parthenon/static_refinementN> blocks written into $WORK (never into $REF).
run_burgers_static () { # name nx nb nscal recon nlim numlevel "regions" extra...
  local name=$1 nx=$2 nb=$3 nscal=$4 recon=$5 nlim=$6 numlevel=$7 regions=$8; shift 8
  local d="$WORK/$name"; rm -rf "$d"; mkdir -p "$d"; cd "$d"
  cp "$REF/benchmarks/burgers/burgers.pin" deck.pin
  local n=0
  for r in $regions; do # level:x1min:x1max:x2min:x2max:x3min:x3max
    IFS=: read -r lev a b c e f g <<< "$r"
    printf '\n<parthenon/static_refinement%d>\nlevel = %s\nx1min = %s\nx1max = %s\nx2min = %s\nx2max = %s\nx3min = %s\nx3max = %s\n' \
      $n $lev $a $b $c $e $f $g >> deck.pin
    n=$((n+1))
  done
  PB2_DUMP_PREFIX="$d/U" "$WORK/burgers_dump" -i deck.pin \
    parthenon/mesh/nx1=$nx parthenon/mesh/nx2=$nx parthenon/mesh/nx3=$nx \
    parthenon/meshblock/nx1=$nb parthenon/meshblock/nx2=$nb parthenon/meshblock/nx3=$nb \
    parthenon/mesh/refinement=static parthenon/mesh/numlevel=$numlevel \
    parthenon/time/nlim=$nlim parthenon/time/tlim=1e9 parthenon/output0/dt=-1 \
    parthenon/output1/dt=1e-9 burgers/num_scalars=$nscal burgers/recon=$recon "$@" \
    > run.log 2>&1
  python3 "$HERE/pack_dumps.py" "$d" "$OUT/$name.npz"
  cp "$d/burgers.out1.hst" "$OUT/$name.hst"
}
if [ -z "${SKIP_STATIC:-}" ]; then
run_burgers_static burgers_s16_b8_l2_weno5 16 8 1 weno5 2 2 "1:0.05:0.2:0.05:0.2:0.05:0.2"
# 2-D, three levels (x3 extents of the regions are ignored in 2-D)
run_burgers_static burgers_s64_b8_l3_2d_weno5 64 8 1 weno5 2 3 \
  "1:-0.3:0.1:-0.2:0.2:0:0 2:-0.12:-0.05:0.02:0.12:0:0" \
  parthenon/mesh/nx3=1 parthenon/meshblock/nx3=1
fi
# example/advection (constant velocity, donor cell): the generic task list with
# AddFluxCorrectionTasks + AddBoundaryExchangeTasks, i.e. prolongation INSIDE every cycle.
run_advection () { # name ndim nx nb nlim refinement numlevel "regions" extra...
  local name=$1 ndim=$2 nx=$3 nb=$4 nlim=$5 refinement=$6 numlevel=$7 regions=$8; shift 8
  local d="$WORK/$name"; rm -rf "$d"; mkdir -p "$d"; cd "$d"
  cp "$REF/example/advection/parthinput.advection" deck.pin
  local n=0
  for r in $regions; do
    IFS=: read -r lev a b c e f g <<< "$r"
    printf '\n<parthenon/static_refinement%d>\nlevel = %s\nx1min = %s\nx1max = %s\nx2min = %s\nx2max = %s\nx3min = %s\nx3max = %s\n' \
      $n $lev $a $b $c $e $f $g >> deck.pin
    n=$((n+1))
  done
  local nx3=1 nb3=1
  if [ "$ndim" = 3 ]; then nx3=$nx; nb3=$nb; fi
  PB2_DUMP_PREFIX="$d/U" PB2_DUMP_FIELD=advected "$WORK/advection_dump" -i deck.pin \
    parthenon/mesh/nx1=$nx parthenon/mesh/nx2=$nx parthenon/mesh/nx3=$nx3 \
    parthenon/meshblock/nx1=$nb parthenon/meshblock/nx2=$nb parthenon/meshblock/nx3=$nb3 \
    parthenon/mesh/refinement=$refinement parthenon/mesh/numlevel=$numlevel \
    parthenon/time/nlim=$nlim parthenon/time/tlim=1e9 parthenon/output0/dt=-1 \
    parthenon/output1/dt=-1 parthenon/output3/dt=-1 parthenon/output4/dt=-1 \
    Advection/fill_derived=false "$@" > run.log 2>&1
  python3 "$HERE/pack_dumps.py" "$d" "$OUT/$name.npz"
}
if [ -z "${SKIP_ADVECTION:-}" ]; then
# BASELINE.json configs[0]: 2-D 256^2 mesh, 32^2 blocks, uniform, hard sphere
run_advection advection_u256_b32_2d_hard_sphere 2 256 32 3 none 1 ""
# 3-D, three static levels (configs[2] without the remesh step), smooth data so that the
# min-mod slopes of the in-cycle prolongation are non-trivial, and the hard sphere of the deck
run_advection advection_s16_b8_l3_gaussian 3 16 8 3 static 3 \
  "1:-0.3:0.2:-0.2:0.3:-0.3:0.1 2:-0.1:0.05:-0.05:0.12:-0.12:0.0" \
  Advection/profile=smooth_gaussian Advection/amp=1.0 Advection/vy=-0.7 Advection/vz=0.4
run_advection advection_s16_b8_l3_hard_sphere 3 16 8 2 static 3 \
  "1:-0.3:0.2:-0.2:0.3:-0.3:0.1 2:-0.1:0.05:-0.05:0.12:-0.12:0.0"
# multilevel AND non-periodic: physical boundaries are applied to the coarse buffers before the
# prolongation and to the fine arrays after it (mesh.cpp:698-706, boundary_communication.cpp:437-449)
run_advection advection_s16_b8_l3_bc 3 16 8 3 static 3 \
  "1:-0.5:0.2:-0.2:0.5:-0.3:0.1 2:-0.5:-0.3:0.3:0.5:-0.12:0.0" \
  Advection/profile=smooth_gaussian Advection/amp=1.0 Advection/vy=-0.7 Advection/vz=0.4 \
  parthenon/mesh/ix1_bc=outflow parthenon/mesh/ox1_bc=outflow \
  parthenon/mesh/ix2_bc=reflecting parthenon/mesh/ox2_bc=reflecting
# ADAPTIVE meshes (configs[2] in small): refinement tagging, tree update with proper nesting,
# refine / derefine data movement, every cycle.  The block list changes => per-cycle metadata.
# derefine_count is lowered so that blocks are also merged within the run.
export PB2_PER_CYCLE_META=1
run_advection advection_a32_b8_l3_2d 2 32 8 40 adaptive 3 "" parthenon/mesh/derefine_count=3
run_advection advection_a32_b8_l2_3d 3 32 8 8 adaptive 2 "" parthenon/mesh/derefine_count=2
unset PB2_PER_CYCLE_META
# 3-D, three levels, 30 cycles (848 -> 764 -> 1198 blocks): too large to commit as raw fields
# (380 MB), so the fixture holds the block list and per-block CRC-32 / interior sum of every cycle
d="$WORK/advection_a32_b8_l3_3d"; rm -rf "$d"; mkdir -p "$d"; cd "$d"
cp "$REF/example/advection/parthinput.advection" deck.pin
PB2_DUMP_PREFIX="$d/U" PB2_DUMP_FIELD=advected "$WORK/advection_dump" -i deck.pin \
  parthenon/mesh/nx1=32 parthenon/mesh/nx2=32 parthenon/mesh/nx3=32 \
  parthenon/meshblock/nx1=8 parthenon/meshblock/nx2=8 parthenon/meshblock/nx3=8 \
  parthenon/mesh/refinement=adaptive parthenon/mesh/numlevel=3 parthenon/mesh/derefine_count=3 \
  parthenon/time/nlim=30 parthenon/time/tlim=1e9 parthenon/output0/dt=-1 parthenon/output1/dt=-1 \
  parthenon/output3/dt=-1 parthenon/output4/dt=-1 Advection/fill_derived=false > run.log 2>&1
python3 "$HERE/pack_checksums.py" "$d" "$OUT/advection_a32_b8_l3_3d_crc.npz" 2
# benchmarks/burgers as shipped (adaptive, <parthenon/refinement0> derivative_order_1 on U(3)),
# small: 32^3 base, 8^3 blocks, tolerances lowered so that the mesh changes within 24 cycles
d="$WORK/burgers_a32_b8_l2"; rm -rf "$d"; mkdir -p "$d"; cd "$d"
PB2_DUMP_PREFIX="$d/U" "$WORK/burgers_dump" -i "$REF/benchmarks/burgers/burgers.pin" \
  parthenon/mesh/nx1=32 parthenon/mesh/nx2=32 parthenon/mesh/nx3=32 \
  parthenon/meshblock/nx1=8 parthenon/meshblock/nx2=8 parthenon/meshblock/nx3=8 \
  parthenon/time/nlim=24 parthenon/output0/dt=-1 parthenon/output1/dt=-1 burgers/num_scalars=1 \
  parthenon/mesh/derefine_count=3 parthenon/refinement0/refine_tol=0.3 \
  parthenon/refinement0/derefine_tol=0.1 > run.log 2>&1
python3 "$HERE/pack_checksums.py" "$d" "$OUT/burgers_a32_b8_l2_crc.npz" 4
fi
# example/sparse_advection (2-D only in the reference): four sparse fields allocated where
# their data is, allocated on a neighbour when a non-null boundary buffer arrives, deallocated
# after dealloc_count quiet cycles.  Uniform mesh; field not allocated on a block => NaN.
run_sparse () { # name nx nb nlim extra...
  local name=$1 nx=$2 nb=$3 nlim=$4; shift 4
  local d="$WORK/$name"; rm -rf "$d"; mkdir -p "$d"; cd "$d"
  PB2_DUMP_PREFIX="$d/U" "$WORK/sparse_dump" -i "$REF/example/sparse_advection/parthinput.sparse_advection" \
    parthenon/mesh/nx1=$nx parthenon/mesh/nx2=$nx parthenon/meshblock/nx1=$nb parthenon/meshblock/nx2=$nb \
    parthenon/mesh/refinement=none parthenon/mesh/numlevel=1 \
    parthenon/time/nlim=$nlim parthenon/time/tlim=1e9 parthenon/output0/dt=-1 \
    parthenon/output1/dt=-1 parthenon/output3/dt=-1 "$@" > run.log 2>&1
  python3 "$HERE/pack_dumps.py" "$d" "$OUT/$name.npz"
}
if [ -z "${SKIP_SPARSE:-}" ]; then
run_sparse sparse_u64_b8_2d 64 8 12
# larger thresholds and a short quiet count so that blocks are DEallocated within the run
PB2_DUMP_EVERY=4 run_sparse sparse_u64_b8_2d_dealloc 64 8 60 parthenon/sparse/alloc_threshold=1e-2 \
  parthenon/sparse/dealloc_threshold=5e-3 parthenon/sparse/dealloc_count=2
fi
# non-cell-centred fields (face / edge / node) after the boundary exchange of Mesh::Initialize:
# keys U_0 (face, 3 elements x 2 components), U_1 (edge, 3 x 1), U_2 (node)
run_tecomm () { # name ndim nx nb ng [numlevel "regions"]
  local name=$1 ndim=$2 nx=$3 nb=$4 ng=$5 numlevel=${6:-1} regions=${7:-}
  local d="$WORK/$name"; rm -rf "$d"; mkdir -p "$d"; cd "$d"
  local nx3=$nx nb3=$nb; if [ "$ndim" = 2 ]; then nx3=1; nb3=1; fi
  local refinement=none; if [ "$numlevel" -gt 1 ]; then refinement=static; fi
  printf '<parthenon/job>\nproblem_id = tecomm\n<parthenon/mesh>\nrefinement = %s\nnumlevel = %d\nnghost = %d\n' $refinement $numlevel $ng > deck.pin
  printf 'nx1 = %d\nx1min = -0.5\nx1max = 0.5\nix1_bc = periodic\nox1_bc = periodic\n' $nx >> deck.pin
  printf 'nx2 = %d\nx2min = -0.5\nx2max = 0.5\nix2_bc = periodic\nox2_bc = periodic\n' $nx >> deck.pin
  printf 'nx3 = %d\nx3min = -0.5\nx3max = 0.5\nix3_bc = periodic\nox3_bc = periodic\n' $nx3 >> deck.pin
  printf '<parthenon/meshblock>\nnx1 = %d\nnx2 = %d\nnx3 = %d\n<parthenon/time>\ntlim = 1.0\nnlim = 0\n' $nb $nb $nb3 >> deck.pin
  local n=0
  for r in $regions; do # level:x1min:x1max:x2min:x2max:x3min:x3max
    IFS=: read -r lev a b c e f g <<< "$r"
    printf '\n<parthenon/static_refinement%d>\nlevel = %s\nx1min = %s\nx1max = %s\nx2min = %s\nx2max = %s\nx3min = %s\nx3max = %s\n' \
      $n $lev $a $b $c $e $f $g >> deck.pin
    n=$((n+1))
  done
  PB2_DUMP_PREFIX="$d/U" "${TECOMM_BIN:-$WORK/tecomm_dump}" -i deck.pin ${TECOMM_ARGS:-} > run.log 2>&1
  python3 "$HERE/pack_dumps.py" "$d" "$OUT/$name.npz"
}
if [ -z "${SKIP_TECOMM:-}" ]; then
run_tecomm tecomm_u16_b8_g2_3d 3 16 8 2
run_tecomm tecomm_u16_b8_g4_3d 3 16 8 4
run_tecomm tecomm_u16_b4_g2_3d 3 16 4 2
run_tecomm tecomm_u32_b8_g2_2d 2 32 8 2
# statically refined meshes: restriction, shared and internal prolongation of face / edge /
# node fields (pins the oracle; the GPU path for these is not built yet)
run_tecomm tecomm_s16_b8_l2_3d 3 16 8 2 2 "1:0.05:0.2:0.05:0.2:0.05:0.2"
run_tecomm tecomm_s32_b8_l3_2d 2 32 8 2 3 "1:-0.3:0.1:-0.2:0.2:0:0 2:-0.12:-0.05:0.02:0.12:0:0"
# nghost = 4 on three levels in 2-D; three levels in 3-D (197 blocks of 4^3: kept as one CRC-32
# per block and field, the convention of pack_checksums.py)
run_tecomm tecomm_s32_b8_g4_l3_2d 2 32 8 4 3 "1:-0.3:0.1:-0.2:0.2:0:0 2:-0.12:-0.05:0.02:0.12:0:0"
run_tecomm tecomm_s16_b4_l3_3d 3 16 4 2 3 "1:-0.3:0.1:-0.2:0.2:-0.1:0.3 2:-0.12:-0.05:0.02:0.12:0.05:0.12"
python3 - "$OUT/tecomm_s16_b4_l3_3d.npz" "$OUT/tecomm_s16_b4_l3_3d_crc.npz" <<'PYEOF'
import sys, zlib
import numpy as np
g = np.load(sys.argv[1])
out = {"meta": g["meta"], "bounds": g["bounds"]}
for k in ("0", "1", "2"):
    a = g["U_" + k]
    out["crc_" + k] = np.array([zlib.crc32(np.ascontiguousarray(a[b]).tobytes())
                                for b in range(a.shape[0])], dtype=np.uint32)
    out["shape_" + k] = np.array(a.shape)
np.savez_compressed(sys.argv[2], **out)
PYEOF
rm -f "$OUT/tecomm_s16_b4_l3_3d.npz"
# the other stock shared prolongations
PB2_SHARED_OP=linear run_tecomm tecomm_s32_b8_l3_2d_linear 2 32 8 2 3 "1:-0.3:0.1:-0.2:0.2:0:0 2:-0.12:-0.05:0.02:0.12:0:0"
PB2_SHARED_OP=constant run_tecomm tecomm_s32_b8_l3_2d_constant 2 32 8 2 3 "1:-0.3:0.1:-0.2:0.2:0:0 2:-0.12:-0.05:0.02:0.12:0:0"
# the face field with ProlongateInternalTothAndRoe (divergence-preserving internal faces)
PB2_TOTH_ROE=1 run_tecomm tecomm_s16_b8_l2_3d_tothroe 3 16 8 2 2 "1:0.05:0.2:0.05:0.2:0.05:0.2"
PB2_TOTH_ROE=1 run_tecomm tecomm_s32_b8_l3_2d_tothroe 2 32 8 2 3 "1:-0.3:0.1:-0.2:0.2:0:0 2:-0.12:-0.05:0.02:0.12:0:0"
# physical boundaries of face / edge / node fields: outflow in x1, reflecting in x2, periodic in
# x3, on a uniform and on the statically refined mesh (coarse-buffer boundaries before the
# prolongation)
BCARGS="parthenon/mesh/ix1_bc=outflow parthenon/mesh/ox1_bc=outflow parthenon/mesh/ix2_bc=reflecting parthenon/mesh/ox2_bc=reflecting"
TECOMM_ARGS="$BCARGS" run_tecomm tecomm_u16_b8_g2_3d_bc 3 16 8 2
TECOMM_ARGS="$BCARGS" run_tecomm tecomm_s16_b8_l2_3d_bc 3 16 8 2 2 "1:0.05:0.5:0.05:0.5:0.05:0.2"
TECOMM_ARGS="$BCARGS" run_tecomm tecomm_s32_b8_l3_2d_bc 2 32 8 2 3 "1:-0.5:0.1:-0.2:0.5:0:0 2:-0.5:-0.3:0.3:0.5:0:0"
# only the face field differs from the fixtures above: drop the edge and node arrays
for n in tecomm_s16_b8_l2_3d_tothroe tecomm_s32_b8_l3_2d_tothroe; do
  python3 -c "import numpy as np,sys; g=np.load(sys.argv[1]); np.savez_compressed(sys.argv[1], **{k: g[k] for k in ('U_0','meta','bounds')})" "$OUT/$n.npz"
done
fi
# adaptive meshes with face / edge / node fields (teamr_dump_main.cpp: the fields never evolve, a
# moving geometric criterion refines and derefines every cycle): remesh data movement of
# non-cell-centred fields, kept as CRC-32 per block, field and cycle
if [ -z "${SKIP_TEAMR:-}" ]; then
if [ ! -x "$WORK/teamr_dump" ] || [ "$HERE/teamr_dump_main.cpp" -nt "$WORK/teamr_dump" ]; then
  $CXX $FLAGS $INC "$HERE/teamr_dump_main.cpp" $LIBS -o "$WORK/teamr_dump"
fi
run_teamr () { # name ndim nx nb numlevel ncycles
  local name=$1 ndim=$2 nx=$3 nb=$4 numlevel=$5 ncyc=$6
  local d="$WORK/$name"; rm -rf "$d"; mkdir -p "$d"; cd "$d"
  local nx3=$nx nb3=$nb; if [ "$ndim" = 2 ]; then nx3=1; nb3=1; fi
  printf '<parthenon/job>\nproblem_id = teamr\n<parthenon/mesh>\nrefinement = adaptive\nnumlevel = %d\nnghost = 2\nderefine_count = 2\n' $numlevel > deck.pin
  printf 'nx1 = %d\nx1min = -0.5\nx1max = 0.5\nix1_bc = periodic\nox1_bc = periodic\n' $nx >> deck.pin
  printf 'nx2 = %d\nx2min = -0.5\nx2max = 0.5\nix2_bc = periodic\nox2_bc = periodic\n' $nx >> deck.pin
  printf 'nx3 = %d\nx3min = -0.5\nx3max = 0.5\nix3_bc = periodic\nox3_bc = periodic\n' $nx3 >> deck.pin
  printf '<parthenon/meshblock>\nnx1 = %d\nnx2 = %d\nnx3 = %d\n<parthenon/time>\ntlim = 1.0\nnlim = 0\n' $nb $nb $nb3 >> deck.pin
  PB2_CYCLES=$ncyc PB2_DUMP_PREFIX="$d/U" "$WORK/teamr_dump" -i deck.pin > run.log 2>&1
  (cd "$HERE" && python3 pack_teamr.py "$d" "$OUT/$name.npz")
}
run_teamr teamr_a32_b8_l3_2d_crc 2 32 8 3 6
run_teamr teamr_a16_b4_l2_3d_crc 3 16 4 2 4
# the same with ProlongateInternalTothAndRoe registered for the face field
PB2_TOTH_ROE=1 run_teamr teamr_a32_b8_l3_2d_tothroe_crc 2 32 8 3 6
PB2_TOTH_ROE=1 run_teamr teamr_a16_b4_l2_3d_tothroe_crc 3 16 4 2 4
fi
# sparse fields on a statically refined mesh (three levels; refined regions on the blobs' paths):
# allocation-aware restriction / prolongation and flux correction
if [ -z "${SKIP_SPARSE:-}" ]; then
  d="$WORK/sparse_s64_b8_l3_2d"; rm -rf "$d"; mkdir -p "$d"; cd "$d"
  cp "$REF/example/sparse_advection/parthinput.sparse_advection" deck.pin
  printf '\n<parthenon/static_refinement0>\nlevel = 1\nx1min = 0.3\nx1max = 0.7\nx2min = 0.3\nx2max = 0.7\n<parthenon/static_refinement1>\nlevel = 2\nx1min = -0.6\nx1max = -0.45\nx2min = 0.45\nx2max = 0.6\n' >> deck.pin
  PB2_DUMP_EVERY=6 PB2_DUMP_PREFIX="$d/U" "$WORK/sparse_dump" -i deck.pin parthenon/mesh/nx1=64 parthenon/mesh/nx2=64 \
    parthenon/meshblock/nx1=8 parthenon/meshblock/nx2=8 parthenon/mesh/refinement=static parthenon/mesh/numlevel=3 \
    parthenon/time/nlim=48 parthenon/time/tlim=1e9 parthenon/output0/dt=-1 parthenon/output1/dt=-1 parthenon/output3/dt=-1 \
    parthenon/sparse/alloc_threshold=1e-2 parthenon/sparse/dealloc_threshold=5e-3 parthenon/sparse/dealloc_count=2 > run.log 2>&1
  python3 "$HERE/pack_dumps.py" "$d" "$OUT/sparse_s64_b8_l3_2d.npz"
fi
# the sparse_advection deck as shipped: refinement = adaptive (three levels), dumps every 5 cycles
# with the block list of every dump
if [ -z "${SKIP_SPARSE:-}" ]; then
  d="$WORK/sparse_a64_b8_l3_2d"; rm -rf "$d"; mkdir -p "$d"; cd "$d"
  PB2_DUMP_EVERY=5 PB2_DUMP_PREFIX="$d/U" "$WORK/sparse_dump" -i "$REF/example/sparse_advection/parthinput.sparse_advection" \
    parthenon/mesh/nx1=64 parthenon/mesh/nx2=64 parthenon/meshblock/nx1=8 parthenon/meshblock/nx2=8 \
    parthenon/mesh/refinement=adaptive parthenon/mesh/numlevel=3 parthenon/time/nlim=40 parthenon/time/tlim=1e9 \
    parthenon/output0/dt=-1 parthenon/output1/dt=-1 parthenon/output3/dt=-1 parthenon/sparse/alloc_threshold=1e-2 \
    parthenon/sparse/dealloc_threshold=5e-3 parthenon/sparse/dealloc_count=2 > run.log 2>&1
  PB2_PER_CYCLE_META=1 python3 "$HERE/pack_dumps.py" "$d" "$OUT/sparse_a64_b8_l3_2d.npz"
fi
# history-only cases (MS Mass 0..7 per cycle, %.14e) at benchmark component count
HST_ONLY=1
run_hst () { local name=$1 nx=$2 nb=$3 nlim=$4
  local d="$WORK/$name"; rm -rf "$d"; mkdir -p "$d"; cd "$d"
  PB2_DUMP_PREFIX="$d/U" PB2_NO_DUMP=1 "$REF_BUILD/benchmarks/burgers/burgers-benchmark" \
    -i "$REF/benchmarks/burgers/burgers.pin" \
    parthenon/mesh/nx1=$nx parthenon/mesh/nx2=$nx parthenon/mesh/nx3=$nx \
    parthenon/meshblock/nx1=$nb parthenon/meshblock/nx2=$nb parthenon/meshblock/nx3=$nb \
    parthenon/mesh/refinement=none parthenon/mesh/numlevel=1 \
    parthenon/time/nlim=$nlim parthenon/time/tlim=1e9 parthenon/output0/dt=-1 \
    parthenon/output1/dt=1e-9 > run.log 2>&1
  cp "$d/burgers.out1.hst" "$OUT/$name.hst"
}
run_hst burgers_u64_b32_s8_weno5 64 32 10
echo "fixtures written to $OUT"

# flux correction of a FACE field (its flux is an edge field): restricted edge fluxes cross
# fine-coarse faces and block edges, the owner's values land (teflux_dump_main.cpp)
if [ ! -x "$WORK/teflux_dump" ] || [ "$HERE/teflux_dump_main.cpp" -nt "$WORK/teflux_dump" ]; then
  $CXX $FLAGS $INC "$HERE/teflux_dump_main.cpp" $LIBS -o "$WORK/teflux_dump"
fi
run_teflux () { # name ndim nx nb ng numlevel "regions"
  local name=$1
  TECOMM_BIN="$WORK/teflux_dump" run_tecomm "$@"
}
if [ -z "${SKIP_TEFLUX:-}" ]; then
run_teflux teflux_s16_b8_l2_3d 3 16 8 2 2 "1:0.05:0.2:0.05:0.2:0.05:0.2"
run_teflux teflux_s32_b8_l3_2d 2 32 8 2 3 "1:-0.3:0.1:-0.2:0.2:0:0 2:-0.12:-0.05:0.02:0.12:0:0"
run_teflux teflux_s16_b4_g4_l3_3d 3 16 4 4 3 "1:-0.3:0.1:-0.2:0.2:-0.1:0.3 2:-0.12:-0.05:0.02:0.12:0.05:0.12"
# 197 blocks: kept as the entries the flux correction changed (flat index into U_0 and value;
# every other entry still holds the generator's code).  Not a CRC: a handful of entries are
# delivered twice and the reference keeps either value (see oracle/pb2_oracle.c)
python3 - "$OUT/teflux_s16_b4_g4_l3_3d.npz" "$OUT/teflux_s16_b4_g4_l3_3d_sparse.npz" <<'PYEOF'
import sys
import numpy as np
g = np.load(sys.argv[1])
a = g["U_0"]
nb, nc, nk, nj, ni = a.shape
init = ((np.arange(nb).reshape(-1, 1, 1, 1, 1) + 1) * 1.0e6 + np.arange(3).reshape(1, -1, 1, 1, 1) * 1.0e5
        + np.arange(nk * nj * ni).reshape(1, 1, nk, nj, ni)).astype(np.float64)
ch = np.flatnonzero(a != init)
np.savez_compressed(sys.argv[2], meta=g["meta"], bounds=g["bounds"], shape_0=np.array(a.shape),
                    changed_idx=ch.astype(np.int64), changed_val=a.reshape(-1)[ch])
PYEOF
rm -f "$OUT/teflux_s16_b4_g4_l3_3d.npz"
fi

# forests of differently oriented trees (ForestDefinition as in example/boundary_exchange): the
# ghost exchange through LogicalCoordinateTransformations; variants in forest_dump_main.cpp
run_forest () { # name variant nb ng
  local name=$1 variant=$2 nb=$3 ng=$4
  local d="$WORK/$name"; rm -rf "$d"; mkdir -p "$d"; cd "$d"
  printf '<parthenon/job>\nproblem_id = forest\n<parthenon/mesh>\nrefinement = static\nnumlevel = 1\nnghost = %d\n<parthenon/meshblock>\nnx1 = %d\nnx2 = %d\nnx3 = 1\n' $ng $nb $nb > deck.pin
  PB2_FOREST_VARIANT=$variant PB2_DUMP_PREFIX="$d/U" "$WORK/forest_dump" -i deck.pin > run.log 2>&1
  python3 "$HERE/pack_dumps.py" "$d" "$OUT/$name.npz"
}
if [ -z "${SKIP_FOREST:-}" ]; then
run_forest forest_v0_b4_g2 0 4 2
run_forest forest_v1_b4_g2 1 4 2
run_forest forest_v2_b8_g2 2 8 2
run_forest forest_v3_b8_g4 3 8 4
fi
