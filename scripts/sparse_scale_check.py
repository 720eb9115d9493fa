#!/usr/bin/env python
"""BASELINE.json configs[3] at its named size on one B200: example/sparse_advection in 3-D,
256^3 mesh of 32^3 blocks (512 blocks), four sparse fields that are allocated where a blob is and
deallocated behind it.  The reference itself aborts in 3-D (sparse_advection_package.cpp:256-257),
so there is NO reference parity at this shape; what is checked here:
  * the full-size run against the CPU oracle (extended with the x3 donor-cell flux) on the first
    cycles where the oracle is still quick: allocation pattern and values bit for bit;
and reported: zone-cycles/s over ALLOCATED (block, field) pairs and over all blocks, and how the
allocated fraction moves."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import oracle  # noqa: E402
from parthenon_b200 import host  # noqa: E402


def state(sim):
    return np.stack([np.where(sim.allocation("base", f"sparse_{f}")[:, None, None, None],
                              sim.get_field("base", f"sparse_{f}")[:, 0], np.nan)
                     for f in range(4)], axis=1)


def main():
    nx, nb = 256, 32
    kw = dict(alloc_threshold=1e-5, dealloc_threshold=1e-6, dealloc_count=5)  # the deck's values
    ov = {"parthenon/mesh/nx1": nx, "parthenon/mesh/nx2": nx, "parthenon/mesh/nx3": nx,
          "parthenon/meshblock/nx1": nb, "parthenon/meshblock/nx2": nb,
          "parthenon/meshblock/nx3": nb}
    sim = host.Simulation(app="sparse_advection", overrides=ov)
    t0 = time.time()
    sim.pre_execute()
    t_init = time.time() - t0
    nblocks = sim.info()["nbtotal"]
    # the first cycles against the oracle (3-D extension, ~10 s of CPU per cycle at this size)
    m = oracle.Mesh(3, (nb,) * 3, 2, (nx // nb,) * 3, xmin=(-1, -1, -1), xmax=(1, 1, 1))
    S = oracle.SparseAdvection(m, **kw)
    S.init()
    ok = sim.dt == S.dt and np.array_equal(state(sim), S.U, equal_nan=True)
    ncheck = 3
    for c in range(ncheck):
        S.step()
        sim.cycle()
        ok = ok and np.array_equal(state(sim), S.U, equal_nan=True)
    print(f"sparse_advection 3-D {nx}^3 / {nb}^3 blocks ({nblocks} blocks x 4 sparse fields): "
          f"init {t_init:.2f} s; first {ncheck} cycles vs the CPU oracle (3-D extension, no "
          f"reference exists in 3-D): {'bit-exact, allocation included' if ok else 'MISMATCH'}",
          flush=True)
    # timing
    ncyc = 60
    alloc = []
    sim.sync()
    t0 = time.time()
    pairs = 0
    for c in range(ncyc):
        a = sum(int(sim.allocation("base", f"sparse_{f}").sum()) for f in range(4))
        alloc.append(a)
        pairs += a
        sim.cycle()
    sim.sync()
    wall = time.time() - t0
    zones = nb ** 3
    print(f"{ncyc} cycles in {wall:.3f} s: allocated (block, field) pairs {min(alloc)}..{max(alloc)} "
          f"of {4 * nblocks} ({100.0 * min(alloc) / (4 * nblocks):.1f}-"
          f"{100.0 * max(alloc) / (4 * nblocks):.1f} %); "
          f"{pairs * zones / wall:.3e} allocated-field zone-cycles/s, "
          f"{ncyc * nblocks * zones / wall:.3e} mesh zone-cycles/s "
          f"({1e3 * wall / ncyc:.2f} ms per cycle)", flush=True)
    # where the time goes: device time per kernel class over a few more cycles
    from parthenon_b200 import capi
    capi.profile(reset=True)
    capi.profile(enable=True)
    t0 = time.time()
    nprof = 20
    for c in range(nprof):
        sim.cycle()
    sim.sync()
    wallp = time.time() - t0
    capi.profile(enable=False)
    prof = capi.profile()
    prof = {k: {"ms": v[0], "launches": v[1]} for k, v in prof.items()}
    tot = sum(v["ms"] for v in prof.values())
    print(f"profile of {nprof} cycles ({1e3 * wallp / nprof:.2f} ms per cycle with event brackets): "
          f"device time {tot / nprof:.3f} ms per cycle: " +
          ", ".join(f"{k} {v['ms'] / nprof:.3f} ms / {v['launches'] // nprof} launches"
                    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]) if v["launches"]),
          flush=True)
    sim.close()
    if not ok:
        raise SystemExit("sparse scale check FAILED")
    print("sparse scale check OK")


if __name__ == "__main__":
    main()
