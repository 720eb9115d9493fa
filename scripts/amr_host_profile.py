#!/usr/bin/env python
"""Where the wall time of an adaptive cycle goes (configs[2]: example/advection 3-D, 128^3 base,
16^3 blocks, 3 levels): Step (device work + task lists) vs regrid (tagging, tree update,
remesh: PB2_TIME_REMESH=1 prints its own phases to stderr)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PB2_TIME_REMESH"] = "1"

from parthenon_b200 import host  # noqa: E402


def main():
    ov = {"parthenon/mesh/refinement": "adaptive", "parthenon/mesh/numlevel": 3,
          "parthenon/mesh/derefine_count": 10, "Advection/profile": "hard_sphere"}
    for d in (1, 2, 3):
        ov[f"parthenon/mesh/nx{d}"] = 128
        ov[f"parthenon/meshblock/nx{d}"] = 16
    sim = host.Simulation(app="advection", overrides=ov)
    sim.pre_execute()
    for _ in range(3):
        sim.cycle()
    sim.sync()
    ts, tr = 0.0, 0.0
    n = 10
    for _ in range(n):
        t0 = time.perf_counter()
        sim.step()
        sim.sync()
        t1 = time.perf_counter()
        sim.regrid()
        sim.sync()
        t2 = time.perf_counter()
        ts += t1 - t0
        tr += t2 - t1
    print(f"blocks {sim.info()['nbtotal']}: step {1e3 * ts / n:.2f} ms, regrid {1e3 * tr / n:.2f} ms "
          f"per cycle")


if __name__ == "__main__":
    main()
