"""Forests of differently oriented trees (ForestDefinition, example/boundary_exchange): the host
framework's forest topology and exchange plan on the CPU against the oracle and the reference's
dumps, and the device exchange — pack, unpack through the neighbour tree's
LogicalCoordinateTransformation, restriction, piecewise-constant prolongation, outflow on the
outer edges — bit for bit against the same dumps (tests/golden/forest_*.npz, made by running the
unmodified reference: tests/golden/refgen/forest_dump_main.cpp)."""
import os

import numpy as np
import pytest

from oracle import forest as F
from parthenon_b200 import host
from tests.test_oracle_golden import FOREST, forest_initial

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def overrides(variant, nb, ng, extra=None):
    ov = {"parthenon/mesh/nghost": ng, "parthenon/meshblock/nx1": nb, "parthenon/meshblock/nx2": nb,
          "forest/variant": variant}
    ov.update(extra or {})
    return ov


@pytest.mark.parametrize("name,variant,nb,ng", FOREST)
def test_host_forest_topology_vs_oracle(name, variant, nb, ng):
    """block list (tree, level, location) in gid order as the reference dumped it; every block's
    neighbour list — gid, level, offsets, in the reference's order — and boundary flags equal the
    oracle's restatement of forest_topology.cpp / tree.cpp"""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    fo = F.lattice_2x2(variant)
    fm = F.ForestMesh(fo, (nb, nb), ng)
    t = host.ForestTopology(variant, overrides(variant, nb, ng))
    info = t.info()
    assert info["nbtotal"] == len(fo.loclist) == len(g["meta"])
    for b, m in enumerate(g["meta"]):
        blk = t.block(b)
        assert blk["gid"] == int(m[0])
        assert blk["loc"] == (int(m[1]), int(m[2]), int(m[3]), int(m[4]))
        assert np.array_equal(blk["xmin"][:2], g["bounds"][b][:2])
        assert np.array_equal(blk["xmax"][:2], g["bounds"][b][3:5])
        want = [(fo.gid[gl], gl[1]) + tuple(F.same_level_offsets(fo.loclist[b], ol))
                for (gl, ol, ct) in fm.neighbors[b]]
        got = [n[:5] for n in t.neighbors(b)]
        assert got == want, b
        flags = t.block_bcs(b)
        for f in range(4):  # user edges take the deck's default, outflow (2); others block (-1)
            assert flags[f] == (2 if fm.block_bc[b, f] else -1), (b, f)
    t.close()


def apply_plan(t, U, ng):
    """the cell-centred exchange of a UNIFORM forest from the host plan alone: same-device fused
    channels copy box to box; channels from differently oriented trees pack the sender's box and
    unpack the buffer through the transformation the plan row carries"""
    new = U.copy()
    rows = list(t.plan_boxes(U.shape[1], 0, "local"))
    send = {(int(r[0]), int(r[1]), int(r[2])): r for r in t.plan_boxes(U.shape[1], 0, "send")}
    for r in rows:
        s, d, n = r[6:9], r[9:12], r[12:15]
        new[r[1], :, d[2]:d[2] + n[2], d[1]:d[1] + n[1], d[0]:d[0] + n[0]] = \
            U[r[0], :, s[2]:s[2] + n[2], s[1]:s[1] + n[1], s[0]:s[0] + n[0]]
    nrecv = 0
    for r in t.plan_boxes(U.shape[1], 0, "recv"):
        sr = send[(int(r[0]), int(r[1]), int(r[2]))]
        assert int(sr[15]) == int(r[15])  # same slab offset on both sides
        s, n = sr[6:9], r[12:15]
        buf = U[r[0], :, s[2]:s[2] + n[2], s[1]:s[1] + n[1], s[0]:s[0] + n[0]]
        assert buf.shape[1:] == (n[2], n[1], n[0])
        flags = int(r[17])
        d = r[9:12]
        if flags & 4:
            dirs = [(flags >> (3 + 2 * q)) & 3 for q in range(3)]
            flip = [(flags >> (9 + q)) & 1 for q in range(3)]
            ncell = flags >> 12
            oracle_unpack(new[r[1]], d, n, buf, dirs, flip, ncell)
        else:
            new[r[1], :, d[2]:d[2] + n[2], d[1]:d[1] + n[1], d[0]:d[0] + n[0]] = buf
        nrecv += 1
    assert nrecv == len(send)
    return new


def oracle_unpack(var, s, n, buf, dirs, flip, ncell):
    import oracle
    v = np.ascontiguousarray(var)
    oracle.unpack_box_transformed(v, [int(x) for x in s], [int(x) for x in n],
                                  np.ascontiguousarray(buf).ravel(), dirs, flip, int(ncell), 1.0)
    var[...] = v


def outflow(U, flags, ng):
    """GenericBC outflow on the faces flagged 2, x1 faces first (boundary_conditions.cpp:47-55)"""
    ni = U.shape[-1]
    if flags[0] == 2:
        U[..., :, :ng] = U[..., :, ng:ng + 1]
    if flags[1] == 2:
        U[..., :, ni - ng:] = U[..., :, ni - ng - 1:ni - ng]
    if flags[2] == 2:
        U[..., :ng, :] = U[..., ng:ng + 1, :]
    if flags[3] == 2:
        U[..., ni - ng:, :] = U[..., ni - ng - 1:ni - ng, :]


@pytest.mark.parametrize("name,variant,nb,ng", [f for f in FOREST if f[1] in (1, 2)])
def test_host_forest_plan_applied_in_numpy(name, variant, nb, ng):
    """uniform forests: the host's exchange plan — fused channels between trees of the same
    orientation, slab channels with a transformation between the others — applied in numpy
    reproduces the reference's dump, ghosts included"""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    ref = g["U_0"]
    t = host.ForestTopology(variant, overrides(variant, nb, ng))
    U = apply_plan(t, forest_initial(*ref.shape), ng)
    for b in range(U.shape[0]):
        outflow(U[b], t.block_bcs(b), ng)
    assert any(int(r[17]) & 4 for r in t.plan_boxes(ref.shape[1], 0, "recv"))
    assert np.array_equal(U, ref)
    t.close()


@pytest.mark.gpu
@pytest.mark.parametrize("extra", [None, {"pb2/virtual_ranks": 2},
                                   {"pb2/virtual_ranks": 2, "pb2/peer_push": "true"}])
@pytest.mark.parametrize("name,variant,nb,ng", FOREST)
def test_forest_exchange_bit_exact(name, variant, nb, ng, extra):
    """the device path on the four forests (example/boundary_exchange as shipped among them):
    after Mesh::Initialize, and again from the generator's state through the exchange tasks,
    every value of every block — ghost cells behind rotated and reflected tree boundaries, behind
    fine-coarse boundaries between trees, and behind the outflow edges — equals the reference's"""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    ref = g["U_0"]
    sim = host.Simulation(app="forest", overrides=overrides(variant, nb, ng, extra))
    try:
        got = sim.get_field("base", "neighbor_info")
        assert got.shape == ref.shape
        assert np.array_equal(got, ref)
        sim.set_field("base", "neighbor_info", forest_initial(*ref.shape))
        sim.exchange("base", prolongate=True)
        assert np.array_equal(sim.get_field("base", "neighbor_info"), ref)
        sim.exchange("base", prolongate=True)
        assert np.array_equal(sim.get_field("base", "neighbor_info"), ref)
    finally:
        sim.close()
