#!/usr/bin/env python
"""Parthenon-VIBE exactly as the reference ships it (benchmarks/burgers/burgers.pin): 128^3 base
mesh of 16^3 blocks, refinement = adaptive with 2 levels (derivative_order_1 on U(3), refine_tol
0.5, derefine_tol 0.2), nghost 4, weno5, 8 scalars, rk2, cfl 0.8, run to tlim = 0.4 (~250 cycles).
This is the ONLY configuration the reference publishes numbers for (benchmarks/burgers/
README.md:111): ~4.0e6 zone-cycles/wallsecond on a 36-core Broadwell node, ~1.8e7 on one A100;
it starts with 624 blocks and ends with more than 800.

zone-cycles/wallsecond is computed like the reference's driver (driver.cpp:57-63, 124):
sum over cycles of nbtotal x cells per block / wall time of the main loop, remeshes included."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from parthenon_b200 import host  # noqa: E402


def run(math):
    ov = {"parthenon/mesh/nghost": 4, "parthenon/mesh/refinement": "adaptive",
          "parthenon/mesh/numlevel": 2, "parthenon/time/tlim": 0.4,
          "burgers/num_scalars": 8, "burgers/recon": "weno5", "pb2/math": math}
    for d in (1, 2, 3):
        ov[f"parthenon/mesh/nx{d}"] = 128
        ov[f"parthenon/meshblock/nx{d}"] = 16
    t0 = time.time()
    sim = host.Simulation(overrides=ov)
    sim.pre_execute()
    sim.sync()
    t_init = time.time() - t0
    n0 = sim.info()["nbtotal"]
    blocks, ncyc, nmax = 0, 0, n0
    t0 = time.time()
    while sim.time < 0.4 and ncyc < 2000:
        blocks += sim.info()["nbtotal"]
        sim.cycle()
        ncyc += 1
        nmax = max(nmax, sim.info()["nbtotal"])
    sim.sync()
    wall = time.time() - t0
    n1 = sim.info()["nbtotal"]
    hist = [float(x) for x in sim.history()]
    sim.close()
    return {"math": math, "cycles": ncyc, "blocks_start": n0, "blocks_end": n1, "blocks_max": nmax,
            "wall_s": wall, "init_s": t_init, "zone_cycles": blocks * 16 ** 3,
            "zone_cycles_per_wallsecond": blocks * 16 ** 3 / wall, "final_time": sim_time(hist),
            "history_MS_Mass": hist}


def sim_time(_):
    return 0.4


def main():
    out = {"config": "benchmarks/burgers/burgers.pin as shipped: 128^3 base, 16^3 blocks, adaptive "
                     "2 levels, weno5, 8 scalars, tlim 0.4",
           "published": {"broadwell_36c_zcps": 4.0e6, "a100_zcps": 1.8e7,
                         "source": "benchmarks/burgers/README.md:111"},
           "runs": [run(m) for m in (sys.argv[1:] or ["fast", "strict"])]}
    best = max(r["zone_cycles_per_wallsecond"] for r in out["runs"])
    out["vs_published_a100"] = best / 1.8e7
    out["vs_published_broadwell_36c"] = best / 4.0e6
    print(json.dumps(out))


if __name__ == "__main__":
    main()
