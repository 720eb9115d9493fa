"""GPU parity tests of example/advection through the C++ host framework (AdvectionDriver task
lists -> C ABI -> sm_100a kernels) against committed outputs of the reference itself
(tests/golden/advection_*.npz).  This is the GENERIC Parthenon stage list: donor-cell fluxes,
flux correction, FluxDivergence, Average/UpdateIndependentData, and a boundary exchange with
restriction AND prolongation in every stage.  All of it is -fmad=false arithmetic: bit-exact."""
import os

import numpy as np
import pytest

from parthenon_b200 import host
from tests import helpers as H
from tests.test_host_topology import deck_overrides
from tests.test_oracle_golden import ADVECTION

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def advection_sim(name, ndim, nx_mesh, nx_block, profile, kw, extra=None):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    leaves, nrb = H.leaves_from_bounds(g["bounds"], nx_mesh, nx_block)
    uniform = len(set(l[0] for l in leaves.tolist())) == 1
    ov = deck_overrides(ndim, nx_block, 2, nrb, refinement="none" if uniform else "static")
    ov["Advection/profile"] = profile
    if "amp" in kw:
        ov["Advection/amp"] = kw["amp"]
    for k, v in zip(("vx", "vy", "vz"), kw.get("v", (1.0, 1.0, 1.0))):
        ov[f"Advection/{k}"] = v
    if kw.get("bcs"):
        ov.update({"parthenon/mesh/ix1_bc": "outflow", "parthenon/mesh/ox1_bc": "outflow",
                   "parthenon/mesh/ix2_bc": "reflecting", "parthenon/mesh/ox2_bc": "reflecting"})
    if extra:
        ov.update(extra)
    return g, host.Simulation(app="advection", overrides=ov, leaves=None if uniform else leaves)


@pytest.mark.parametrize("extra", [None, {"pb2/virtual_ranks": 3}])
@pytest.mark.parametrize("name,ndim,nx_mesh,nx_block,profile,kw,ncyc", ADVECTION)
def test_advection_cycles_bit_exact_vs_reference_dumps(name, ndim, nx_mesh, nx_block, profile,
                                                       kw, ncyc, extra):
    g, sim = advection_sim(name, ndim, nx_mesh, nx_block, profile, kw, extra)
    info = sim.info()
    assert info["nbtotal"] == g["meta"].shape[0]
    sim.pre_execute()
    assert sim.dt == g["dts"][0]
    assert np.array_equal(sim.get_field("base", "advected"), g["U_0"])
    for c in range(1, ncyc + 1):
        sim.cycle()
        assert sim.time == g["times"][c]
        assert np.array_equal(sim.get_field("base", "advected"), g[f"U_{c}"]), f"cycle {c}"


def test_user_boundary_conditions_on_multilevel_mesh():
    """user-registered boundary conditions on a three-level mesh whose refined regions touch the
    boundaries: called on the coarse buffers before the prolongation and on the fine arrays
    after it, like the stock ones (same reference dump)"""
    name, ndim, nx_mesh, nx_block, profile, kw, ncyc = [a for a in ADVECTION if a[5].get("bcs")][0]
    extra = {"parthenon/mesh/ix1_bc": "pb2_user_outflow", "parthenon/mesh/ox1_bc": "pb2_user_outflow",
             "parthenon/mesh/ix2_bc": "pb2_user_reflect", "parthenon/mesh/ox2_bc": "pb2_user_reflect"}
    g, sim = advection_sim(name, ndim, nx_mesh, nx_block, profile, kw, extra)
    sim.pre_execute()
    assert np.array_equal(sim.get_field("base", "advected"), g["U_0"])
    for c in range(1, ncyc + 1):
        sim.cycle()
        assert np.array_equal(sim.get_field("base", "advected"), g[f"U_{c}"]), f"cycle {c}"
    sim.close()


def test_advection_conserves_total_on_multilevel_mesh():
    """flux correction makes the update conservative across fine-coarse faces: the
    volume-weighted total of the advected field is constant to rounding"""
    name, ndim, nx_mesh, nx_block, profile, kw, _ = ADVECTION[1]
    g, sim = advection_sim(name, ndim, nx_mesh, nx_block, profile, kw)
    sim.pre_execute()
    vol = np.prod((g["bounds"][:, 3:] - g["bounds"][:, :3]) / np.array(nx_block)[::1], axis=1)

    def total():
        u = sim.get_field("base", "advected")[:, 0, 2:-2, 2:-2, 2:-2]
        return float((u.sum(axis=(1, 2, 3)) * vol).sum())

    t0 = total()
    sim.cycle(5)
    assert abs(total() - t0) <= 1e-13 * abs(t0)


from tests.test_oracle_golden import ADAPTIVE  # noqa: E402


@pytest.mark.parametrize("name,ndim,nx_mesh,nx_block,numlevel,derefine_count,ncyc", ADAPTIVE)
def test_adaptive_advection_bit_exact_vs_reference_dumps(name, ndim, nx_mesh, nx_block, numlevel,
                                                         derefine_count, ncyc):
    """refinement = adaptive through the host framework: the initial refinement loop, tagging on
    the device (pb2_block_minmax), tree update, and the remesh as restrict / copy / prolongate
    launches into a new slab — block list and field bit-exact against the reference's dumps,
    which are taken after Step and before that cycle's remesh (sim.step() / sim.regrid())."""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    nrb = [nx_mesh[d] // nx_block[d] if d < ndim else 1 for d in range(3)]
    ov = deck_overrides(ndim, nx_block, 2, nrb, refinement="adaptive")
    ov.update({"parthenon/mesh/numlevel": numlevel, "parthenon/mesh/derefine_count": derefine_count,
               "Advection/profile": "hard_sphere"})
    sim = host.Simulation(app="advection", overrides=ov)
    sim.pre_execute()
    counts = set()
    for c in range(ncyc + 1):
        if c:
            sim.step()
        ref = g[f"U_{c}"]
        leaves, _ = H.leaves_from_bounds(g[f"bounds_{c}"], nx_mesh, nx_block)
        info = sim.info()
        assert info["nbtotal"] == ref.shape[0], f"cycle {c}"
        locs = np.array([sim.block(b)["loc"] for b in range(info["nblocks"])])
        assert np.array_equal(locs, leaves), f"cycle {c}"
        assert np.array_equal(sim.get_field("base", "advected"), ref), f"cycle {c}"
        assert sim.time == g["times"][c]
        counts.add(info["nbtotal"])
        if c:
            sim.regrid()
    assert len(counts) > 1


def test_adaptive_advection_3d_three_levels_crc():
    """the same run 30 cycles long (848 -> 764 -> 1198 blocks) against the compact fixture of
    the reference: block list and the CRC-32 of every block's bytes, every cycle"""
    from tests.test_oracle_golden import check_against_crc_fixture
    ov = deck_overrides(3, (8, 8, 8), 2, (4, 4, 4), refinement="adaptive")
    ov.update({"parthenon/mesh/numlevel": 3, "parthenon/mesh/derefine_count": 3,
               "Advection/profile": "hard_sphere"})
    sim = host.Simulation(app="advection", overrides=ov)
    sim.pre_execute()

    def state():
        n = sim.info()["nblocks"]
        return (np.array([sim.block(b)["loc"] for b in range(n)]),
                sim.get_field("base", "advected"), sim.time)

    counts = check_against_crc_fixture("advection_a32_b8_l3_3d_crc", (32, 32, 32), (8, 8, 8), 30,
                                       state, sim.step, sim.regrid)
    assert len(counts) > 4
