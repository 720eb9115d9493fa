"""GPU tests of device paths written after the GPU budget of round 1 was spent.  They first ran
on the driver's box at the end of that round (GPUTEST_r01.json: all passed) and are ordinary
tests since: a mismatch fails the suite."""
import os

import numpy as np
import pytest

from parthenon_b200 import host
from tests import helpers as H
from tests.test_host_topology import deck_overrides

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name,ndim,nx,nb,numlevel", H.TEAMR_TOTH_ROE)
def test_adaptive_remesh_with_toth_roe_crc(name, ndim, nx, nb, numlevel):
    """the adaptive runs with ProlongateInternalTothAndRoe registered for the face field
    (tecomm/toth_roe = true): new children get their internal faces from the divergence-
    preserving operator after the newly-refined-ownership exchange"""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    full = lambda n: (n,) * ndim + (1,) * (3 - ndim)
    ov = deck_overrides(ndim, (nb,) * 3, 2, (nx // nb,) * 3, refinement="adaptive")
    ov.update({"parthenon/mesh/numlevel": numlevel, "parthenon/mesh/derefine_count": 2,
               "tecomm/toth_roe": "true"})
    sim = host.Simulation(app="tecomm", overrides=ov)
    try:
        for c in range(int(g["ncycles"]) + 1):
            if c:
                sim.tag_and_remesh(c)
            leaves, _ = H.leaves_from_bounds(g[f"bounds_{c}"], full(nx), full(nb))
            n = sim.info()["nblocks"]
            assert n == len(leaves), c
            assert np.array_equal(np.array([sim.block(b)["loc"] for b in range(n)]), leaves), c
            for f, field in enumerate(("face", "edge", "node")):
                got = sim.get_field("base", field)
                bad = np.nonzero(H.block_crcs(got) != g[f"crc_{c}_{f}"])[0]
                assert len(bad) == 0, (c, field, len(bad), bad[:8])
    finally:
        sim.close()


def test_sparse_advection_on_refined_mesh():
    """sparse fields on a three-level statically refined mesh (same-device channels):
    allocation-aware restriction / prolongation / flux correction vs the reference's dumps"""
    g = np.load(os.path.join(GOLD, "sparse_s64_b8_l3_2d.npz"))
    leaves, nrb = H.leaves_from_bounds(g["bounds"], (64, 64, 1), (8, 8, 1), xmin=-1.0, xmax=1.0)
    ov = {"parthenon/mesh/refinement": "static", "parthenon/mesh/numlevel": 3,
          "parthenon/sparse/alloc_threshold": 1e-2, "parthenon/sparse/dealloc_threshold": 5e-3,
          "parthenon/sparse/dealloc_count": 2}
    sim = host.Simulation(app="sparse_advection", overrides=ov, leaves=leaves)
    try:
        sim.pre_execute()

        def state():
            return np.stack([np.where(sim.allocation("base", f"sparse_{f}")[:, None, None, None],
                                      sim.get_field("base", f"sparse_{f}")[:, 0], np.nan)
                             for f in range(4)], axis=1)

        dumped = {int(c): i for i, c in enumerate(g["cycles"])}
        assert np.array_equal(state(), g["U_0"], equal_nan=True)
        for c in range(1, 49):
            sim.cycle()
            if c in dumped:
                assert np.array_equal(state(), g[f"U_{c}"], equal_nan=True), f"cycle {c}"
    finally:
        sim.close()


def test_sparse_advection_adaptive_as_shipped():
    """example/sparse_advection as the reference ships it — refinement = adaptive, three levels,
    40 cycles (220 -> 328 -> 280 blocks): tagging on the allocated fields only, remesh of sparse
    fields on the device (a new child exists where its parent did, a new parent where any
    daughter did), allocation-aware exchange on the new mesh; block list, allocation pattern and
    every value against the reference's dumps (taken after Step, before that cycle's remesh)"""
    g = np.load(os.path.join(GOLD, "sparse_a64_b8_l3_2d.npz"))
    ov = {"parthenon/mesh/refinement": "adaptive", "parthenon/mesh/numlevel": 3,
          "parthenon/sparse/alloc_threshold": 1e-2, "parthenon/sparse/dealloc_threshold": 5e-3,
          "parthenon/sparse/dealloc_count": 2}
    sim = host.Simulation(app="sparse_advection", overrides=ov)
    try:
        sim.pre_execute()

        def state():
            return np.stack([np.where(sim.allocation("base", f"sparse_{f}")[:, None, None, None],
                                      sim.get_field("base", f"sparse_{f}")[:, 0], np.nan)
                             for f in range(4)], axis=1)

        dumped = {int(c): i for i, c in enumerate(g["cycles"])}
        counts = set()
        for c in range(41):
            if c:
                sim.step()
            if c in dumped:
                leaves, _ = H.leaves_from_bounds(g[f"bounds_{c}"], (64, 64, 1), (8, 8, 1), xmin=-1.0,
                                                 xmax=1.0)
                info = sim.info()
                locs = np.array([sim.block(b)["loc"] for b in range(info["nblocks"])])
                assert info["nbtotal"] == len(leaves) and np.array_equal(locs, leaves), f"cycle {c}"
                assert sim.time == g["times"][dumped[c]]
                assert np.array_equal(state(), g[f"U_{c}"], equal_nan=True), f"cycle {c}"
                counts.add(info["nbtotal"])
            if c:
                sim.regrid()
        assert len(counts) > 3
    finally:
        sim.close()
