#!/usr/bin/env python
"""BASELINE.json configs[2] at its named size on one B200: example/advection 3-D, 128^3 base mesh
of 16^3 blocks, refinement = adaptive with 3 levels, hard sphere.  No reference dump exists at
this size (a dump is ~0.25 GB per cycle), so this script checks what does not need one:
  * a medium case (64^3 base, 16^3 blocks, 3 levels) cycle by cycle against the CPU oracle,
    bit for bit, block lists included;
  * the full-size run: it completes, the block count follows the sphere, and the
    volume-weighted total of the advected field drifts no more than the reference's own runs do
    (the reference is conservative to rounding for ~20 cycles and then drifts at the 1e-12 ..
    1e-10 level on 3-level 3-D meshes — tests/golden/advection_a32_b8_l3_3d_crc.npz was checked
    for this — so the total is reported, with a loose bound);
and reports zone-cycles/s (block-cycles x 16^3 / wall time, remeshes included)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import oracle  # noqa: E402
from parthenon_b200 import host  # noqa: E402
from tests.test_host_topology import deck_overrides  # noqa: E402


def make(nbase, nb, numlevel, dc):
    ov = deck_overrides(3, (nb,) * 3, 2, (nbase // nb,) * 3, refinement="adaptive")
    ov.update({"parthenon/mesh/numlevel": numlevel, "parthenon/mesh/derefine_count": dc,
               "Advection/profile": "hard_sphere"})
    return host.Simulation(app="advection", overrides=ov)


def total(sim, nb):
    info = sim.info()
    u = sim.get_field("base", "advected")[:, 0, 2:-2, 2:-2, 2:-2].sum(axis=(1, 2, 3))
    vol = np.array([np.prod((sim.block(b)["xmax"] - sim.block(b)["xmin"]) / nb)
                    for b in range(info["nblocks"])])
    return float((u * vol).sum()), info["nbtotal"]


def main():
    # medium case against the oracle
    sim = make(64, 16, 3, 3)
    A = oracle.AmrAdvection(3, (16,) * 3, 2, (4,) * 3, 3, derefine_count=3)
    sim.pre_execute()
    A.init()
    ok = True
    for c in range(7):
        if c:
            sim.step()
            A.step()
        locs = np.array([sim.block(b)["loc"] for b in range(sim.info()["nblocks"])])
        ok = ok and np.array_equal(locs, A.block_locs) and \
            np.array_equal(sim.get_field("base", "advected"), A.U) and sim.time == A.time
        if c:
            sim.regrid()
            A.regrid()
    print(f"64^3 base / 16^3 blocks / 3 levels, 6 cycles vs CPU oracle "
          f"({A.nblocks} blocks): {'bit-exact' if ok else 'MISMATCH'}", flush=True)
    sim.close()
    # configs[2] at full size
    sim = make(128, 16, 3, 10)
    t0 = time.time()
    sim.pre_execute()
    t_init = time.time() - t0
    m0, n0 = total(sim, 16)
    counts, blocks = [n0], 0
    t0 = time.time()
    ncyc = 40
    for c in range(ncyc):
        blocks += sim.info()["nbtotal"]
        sim.cycle()
        counts.append(sim.info()["nbtotal"])
    sim.sync()
    wall = time.time() - t0
    m1, n1 = total(sim, 16)
    cons = abs(m1 - m0) / abs(m0)
    print(f"128^3 base / 16^3 blocks / 3 levels: init {t_init:.2f} s -> {n0} blocks; {ncyc} cycles, "
          f"blocks {min(counts)}..{max(counts)} (end {n1}); total advected {m0:.15e} -> {m1:.15e} "
          f"(rel. change {cons:.2e}); {blocks * 16 ** 3 / wall:.3e} zone-cycles/s incl. remesh",
          flush=True)
    if not ok or cons > 1e-8:
        raise SystemExit("AMR scale check FAILED")
    print("AMR scale check OK")


if __name__ == "__main__":
    main()
