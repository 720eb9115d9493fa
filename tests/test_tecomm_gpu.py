"""GPU parity of the ghost exchange of NON-CELL-CENTRED fields (face / edge / node) through the
C++ host framework and the C ABI's copy / pack / unpack kernels, against dumps of the reference
itself (tests/golden/tecomm_*.npz, made by tests/golden/refgen/tecomm_dump_main.cpp): every
entry of the fixtures encodes the block and entry it was copied from, so the test pins the
element-aware index boxes AND which block owns every shared face, edge and node."""
import os

import numpy as np
import pytest

from parthenon_b200 import host
from tests import helpers as H
from tests.test_host_topology import deck_overrides

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
FIELDS = {"face": ("U_0", 3, 2), "edge": ("U_1", 3, 1), "node": ("U_2", 1, 1)}


@pytest.mark.parametrize("extra", [None, {"pb2/virtual_ranks": 2}, {"pb2/virtual_ranks": 3},
                                   {"pb2/virtual_ranks": 3, "pb2/peer_push": "true"}])
@pytest.mark.parametrize("name,ndim,nx,nb,ng", H.TECOMM)
def test_non_cell_centred_exchange_bit_exact(name, ndim, nx, nb, ng, extra):
    """extra = pb2/virtual_ranks splits the blocks of the one GPU into groups that talk through
    the pack -> slab -> unpack path of inter-device channels instead of the fused copy;
    with pb2/peer_push through stores into the receiver's ghost cells + arrival flags"""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    ov = deck_overrides(ndim, (nb,) * 3, ng, (nx // nb,) * 3)
    ov.update(extra or {})
    sim = host.Simulation(app="tecomm", overrides=ov)
    try:
        for field, (key, nel, ncomp) in FIELDS.items():
            ref = g[key]
            # Mesh::Initialize: problem generator, then CommunicateBoundaries
            got = sim.get_field("base", field)
            assert got.shape == ref.shape
            assert np.array_equal(got, ref), (name, field)
        # once more from the generator's state through the exchange tasks, and idempotence
        for field, (key, nel, ncomp) in FIELDS.items():
            ref = g[key]
            nblocks, _, nk, nj, ni = ref.shape
            sim.set_field("base", field,
                          H.tecomm_initial(nblocks, nel, ncomp, nk, nj, ni).reshape(ref.shape))
        sim.exchange("base")
        for field, (key, _, _) in FIELDS.items():
            assert np.array_equal(sim.get_field("base", field), g[key]), (name, field)
        sim.exchange("base")
        for field, (key, _, _) in FIELDS.items():
            assert np.array_equal(sim.get_field("base", field), g[key]), (name, field, "again")
    finally:
        sim.close()


@pytest.mark.parametrize("extra", [None, {"pb2/virtual_ranks": 2}, {"pb2/virtual_ranks": 3}])
@pytest.mark.parametrize("name,ndim,nx,nb,ng", H.TECOMM_MULTILEVEL)
def test_non_cell_centred_multilevel_exchange_bit_exact(name, ndim, nx, nb, ng, extra):
    """statically refined meshes: element forms of RestrictAverage / ProlongateSharedMinMod /
    ProlongateInternalAverage on the device, ownership masks resolved into boxes, against the
    reference's dumps (3-D two levels, 2-D three levels); also with the blocks split into
    virtual ranks so that fine-coarse channels cross the pack / slab / unpack path"""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    full = lambda n: (n,) * ndim + (1,) * (3 - ndim)
    leaves, nrb = H.leaves_from_bounds(g["bounds"], full(nx), full(nb))
    ov = deck_overrides(ndim, (nb,) * 3, ng, nrb, refinement="static")
    ov.update(extra or {})
    sim = host.Simulation(app="tecomm", overrides=ov, leaves=leaves)
    try:
        for field, (key, nel, ncomp) in FIELDS.items():
            got = sim.get_field("base", field)
            assert got.shape == g[key].shape
            assert np.array_equal(got, g[key]), (name, field)
        for field, (key, nel, ncomp) in FIELDS.items():
            ref = g[key]
            nblocks, _, nk, nj, ni = ref.shape
            sim.set_field("base", field,
                          H.tecomm_initial(nblocks, nel, ncomp, nk, nj, ni).reshape(ref.shape))
        sim.exchange("base")
        for field, (key, _, _) in FIELDS.items():
            assert np.array_equal(sim.get_field("base", field), g[key]), (name, field)
    finally:
        sim.close()


@pytest.mark.parametrize("name,ndim,nx,nb,ng,op,opname", H.TECOMM_SHARED_OPS)
def test_non_cell_centred_other_shared_prolongations(name, ndim, nx, nb, ng, op, opname):
    """tecomm/shared_op = linear | constant: ProlongateSharedLinear / ProlongatePiecewiseConstant
    on faces, edges and nodes (pb2_prolongate_te with the operator id) vs the reference"""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    full = lambda n: (n,) * ndim + (1,) * (3 - ndim)
    leaves, nrb = H.leaves_from_bounds(g["bounds"], full(nx), full(nb))
    ov = deck_overrides(ndim, (nb,) * 3, ng, nrb, refinement="static")
    ov["tecomm/shared_op"] = opname
    sim = host.Simulation(app="tecomm", overrides=ov, leaves=leaves)
    try:
        for field, (key, nel, ncomp) in FIELDS.items():
            assert np.array_equal(sim.get_field("base", field), g[key]), (name, field)
    finally:
        sim.close()


@pytest.mark.parametrize("extra", [None, {"pb2/virtual_ranks": 3}])
@pytest.mark.parametrize("name,ndim,nx,nb,ng", H.TECOMM_MULTILEVEL_CRC)
def test_non_cell_centred_three_levels_3d_crc(name, ndim, nx, nb, ng, extra):
    """three levels in 3-D (197 blocks of 4^3) against the CRC-32 per block and field of the
    reference's dump"""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    full = lambda n: (n,) * ndim + (1,) * (3 - ndim)
    leaves, nrb = H.leaves_from_bounds(g["bounds"], full(nx), full(nb))
    ov = deck_overrides(ndim, (nb,) * 3, ng, nrb, refinement="static")
    ov.update(extra or {})
    sim = host.Simulation(app="tecomm", overrides=ov, leaves=leaves)
    try:
        for field, (key, nel, ncomp) in FIELDS.items():
            got = sim.get_field("base", field)
            assert got.shape == tuple(g["shape_" + key[2:]])
            assert np.array_equal(H.block_crcs(got), g["crc_" + key[2:]]), (name, field)
    finally:
        sim.close()


@pytest.mark.parametrize("extra", [None, {"pb2/virtual_ranks": 2}])
@pytest.mark.parametrize("name,ndim,nx,nb,ng", H.TECOMM_TOTH_ROE)
def test_toth_roe_internal_prolongation_bit_exact(name, ndim, nx, nb, ng, extra):
    """the face field registered with ProlongateInternalTothAndRoe (tecomm/toth_roe = true):
    pb2_prolongate_toth_roe against the reference's dumps"""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    full = lambda n: (n,) * ndim + (1,) * (3 - ndim)
    leaves, nrb = H.leaves_from_bounds(g["bounds"], full(nx), full(nb))
    ov = deck_overrides(ndim, (nb,) * 3, ng, nrb, refinement="static")
    ov["tecomm/toth_roe"] = "true"
    ov.update(extra or {})
    sim = host.Simulation(app="tecomm", overrides=ov, leaves=leaves)
    try:
        assert np.array_equal(sim.get_field("base", "face"), g["U_0"]), name
    finally:
        sim.close()


@pytest.mark.parametrize("extra", [None, {"pb2/virtual_ranks": 2}])
@pytest.mark.parametrize("name,ndim,nx,nb,ng", H.TECOMM_BC)
def test_non_cell_centred_physical_boundaries_bit_exact(name, ndim, nx, nb, ng, extra):
    """outflow (x1) / reflecting (x2) mesh boundaries of face / edge / node fields: pb2_apply_bcs
    with one region per topological element, on coarse buffers before the prolongation and on
    the fine arrays after it, against the reference's dumps"""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    full = lambda n: (n,) * ndim + (1,) * (3 - ndim)
    leaves, nrb = H.leaves_from_bounds(g["bounds"], full(nx), full(nb))
    static = len(set(g["meta"][:, 1])) > 1
    ov = deck_overrides(ndim, (nb,) * 3, ng, nrb, refinement="static" if static else "none")
    for key, val in zip(("ix1_bc", "ox1_bc", "ix2_bc", "ox2_bc"), H.TECOMM_BC_NAMES):
        ov["parthenon/mesh/" + key] = val
    ov.update(extra or {})
    sim = host.Simulation(app="tecomm", overrides=ov, leaves=leaves if static else None)
    try:
        for field, (key, nel, ncomp) in FIELDS.items():
            assert np.array_equal(sim.get_field("base", field), g[key]), (name, field)
    finally:
        sim.close()


@pytest.mark.parametrize("name,ndim,nx,nb,numlevel", H.TEAMR)
def test_adaptive_remesh_of_non_cell_centred_fields_crc(name, ndim, nx, nb, numlevel):
    """adaptive meshes: every cycle a moving geometric criterion refines some blocks and derefines
    others while the fields never evolve, so the state after each remesh pins the data movement
    of face / edge / node fields (element-wise restriction into new parents with shared elements
    from the upper daughter, shared prolongation into new children, an exchange in which older
    fine blocks own shared elements over newly refined ones, internal prolongation, the regular
    exchange) — block lists and one CRC-32 per block and field against the reference"""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    full = lambda n: (n,) * ndim + (1,) * (3 - ndim)
    ov = deck_overrides(ndim, (nb,) * 3, 2, (nx // nb,) * 3, refinement="adaptive")
    ov.update({"parthenon/mesh/numlevel": numlevel, "parthenon/mesh/derefine_count": 2})
    sim = host.Simulation(app="tecomm", overrides=ov)
    try:
        counts = set()
        for c in range(int(g["ncycles"]) + 1):
            if c:
                sim.tag_and_remesh(c)
            leaves, _ = H.leaves_from_bounds(g[f"bounds_{c}"], full(nx), full(nb))
            n = sim.info()["nblocks"]
            assert n == len(leaves), c
            assert np.array_equal(np.array([sim.block(b)["loc"] for b in range(n)]), leaves), c
            for f, field in enumerate(("face", "edge", "node")):
                got = sim.get_field("base", field)
                assert got.shape == tuple(g[f"shape_{c}_{f}"])
                bad = np.nonzero(H.block_crcs(got) != g[f"crc_{c}_{f}"])[0]
                assert len(bad) == 0, (c, field, len(bad), bad[:8])
            counts.add(n)
        assert len(counts) > 2
    finally:
        sim.close()


TEFLUX = [("teflux_s16_b8_l2_3d", 3, 16, 8, 2), ("teflux_s32_b8_l3_2d", 2, 32, 8, 2),
          ("teflux_s16_b4_g4_l3_3d_sparse", 3, 16, 4, 4)]


@pytest.mark.parametrize("extra", [None, {"pb2/virtual_ranks": 2}, {"pb2/virtual_ranks": 3}])
@pytest.mark.parametrize("name,ndim,nx,nb,ng", TEFLUX)
def test_flux_correction_of_a_face_field_bit_exact(name, ndim, nx, nb, ng, extra):
    """flux correction of a FACE field on the device: its flux is the edge field "bnd_flux::B"
    (StateDescriptor::AddField).  Fine blocks restrict the edge elements they share with coarser
    neighbours (restrict_te_kernel), the coarser blocks take the entries the sender owns
    (copy_kernel, block-edge messages first, face messages second), against dumps of the
    reference (tests/golden/refgen/teflux_dump_main.cpp: 3-D two levels, 2-D three levels, 3-D
    three levels of 4^3 blocks with 4 ghosts).  Entries the reference delivers twice in a
    shuffled order (see oracle/pb2_oracle.c) must hold the oracle's choice, every other entry
    the reference's; a second correction changes nothing.  extra = pb2/virtual_ranks: fine-coarse
    pairs in different groups of blocks take the inter-device path (restrict, pack what the
    sender owns into the flux-correction slab, unpack block-edge messages before face messages)"""
    import oracle
    from tests.test_oracle_golden import teflux_initial, teflux_reference
    g = np.load(os.path.join(GOLD, name + ".npz"))
    full = lambda n: (n,) * ndim + (1,) * (3 - ndim)
    leaves, nrb = H.leaves_from_bounds(g["bounds"], full(nx), full(nb))
    ref = teflux_reference(g)
    m = oracle.Mesh(ndim, (nb,) * ndim, ng, tuple(nrb[:ndim]), leaves=leaves)
    F = teflux_initial(ref.shape[0], *ref.shape[2:])
    init = F[:, :, 0].copy()
    w = np.zeros(F.shape, dtype=np.int32)
    assert m.flux_correct_edge(F, w) > 0
    once = w[:, :, 0] <= 1
    ov = deck_overrides(ndim, (nb,) * 3, ng, nrb, refinement="static")
    ov["tecomm/flux_field"] = "true"
    ov.update(extra or {})
    sim = host.Simulation(app="tecomm", overrides=ov, leaves=leaves)
    try:
        assert sim.field_shape("base", "bnd_flux::B") == ref.shape
        if extra:
            assert len(sim.edge_flux_plan("send")) > 0
        sim.set_field("base", "bnd_flux::B", init)
        sim.flux_correction("base")
        got = sim.get_field("base", "bnd_flux::B")
        assert np.array_equal(got[once], ref[once]), name
        assert np.array_equal(got, F[:, :, 0]), name
        assert np.array_equal(got != init, ref != init)
        sim.flux_correction("base")
        assert np.array_equal(sim.get_field("base", "bnd_flux::B"), got), (name, "again")
        # the face field itself still exchanges like any other
        sim.exchange("base")
    finally:
        sim.close()
