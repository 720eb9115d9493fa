#!/usr/bin/env python
"""Host half of a remesh without a device (development container): tree update + block list
(Topology.regrid) and the exchange plan of an adaptive mesh of several thousand blocks, with the
phase timers of PB2_TIME_HOST=1 ([pb2 regrid] / [pb2 rebuild] / [pb2 plan] lines on stderr)."""
import time, numpy as np, sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from parthenon_b200 import host
from parthenon_b200.host import lib
import ctypes as C
ov = {"parthenon/mesh/refinement": "adaptive", "parthenon/mesh/numlevel": 3, "parthenon/mesh/nghost": 2}
for d in (1, 2, 3):
    ov[f"parthenon/mesh/nx{d}"] = 128
    ov[f"parthenon/meshblock/nx{d}"] = 16
t0 = time.perf_counter()
t = host.Topology(deck=host.ADVECTION_DECK, overrides=ov)
print("create", time.perf_counter() - t0, t.info()["nbtotal"])
def centers():
    n = t.info()["nbtotal"]
    c = np.zeros((n, 3))
    for b in range(n):
        blk = t.block(b)
        c[b] = 0.5 * (np.array(blk["xmin"]) + np.array(blk["xmax"]))
    return c
def tags_for(shift):
    c = centers()
    r = np.linalg.norm(c - np.array([shift, 0, 0]), axis=1)
    # refine near the sphere surface r = 0.3 (units of the domain -0.5..0.5)
    tg = np.where(np.abs(r - 0.3) < 0.06, 1, -1).astype(np.int32)
    return tg
for it in range(8):
    tg = tags_for(0.0 + 0.01 * max(0, it - 3))
    t0 = time.perf_counter()
    ch = t.regrid(tg)
    t1 = time.perf_counter()
    n = lib().pb2h_sim_plan_boxes(t.h, 1, 0, 0, None, 0)
    t2 = time.perf_counter()
    print(it, "regrid", round(1e3 * (t1 - t0), 2), "ms changed", ch, "blocks", t.info()["nbtotal"],
          "plan", round(1e3 * (t2 - t1), 2), "ms rows", n)
