"""CPU tests of the C++ host framework's topology (no device touched): block order,
coordinates, neighbour lists and boundary index boxes must equal the oracle's (which is pinned
to the reference), rank assignment must follow the reference's AssignBlocks, and the per-peer
slab layouts computed independently by two ranks must agree."""
import os

import numpy as np
import pytest

import oracle
from parthenon_b200 import host
from tests import helpers as H


def deck_overrides(ndim, nx, ng, nrb, refinement="none"):
    ov = {"parthenon/mesh/nghost": ng, "parthenon/mesh/refinement": refinement}
    for d in range(3):
        n = d + 1
        if d < ndim:
            ov[f"parthenon/mesh/nx{n}"] = nx[d] * nrb[d]
            ov[f"parthenon/meshblock/nx{n}"] = nx[d]
        else:
            ov[f"parthenon/mesh/nx{n}"] = 1
            ov[f"parthenon/meshblock/nx{n}"] = 1
    return ov


def compare_topology(t, m):
    info = t.info()
    assert info["nbtotal"] == m.nblocks
    assert (info["nk"], info["nj"], info["ni"]) == m.dims
    if m.multilevel:
        assert (info["cnk"], info["cnj"], info["cni"]) == m.cdims
    for b in range(m.nblocks):
        blk = t.block(b)
        assert blk["loc"] == m.block_loc(b)
        lo, hi = m.block_bounds(b)
        assert np.array_equal(blk["xmin"], lo) and np.array_equal(blk["xmax"], hi)
        nbs = t.neighbors(b)
        ref = m.neighbors(b)
        assert [x[:5] for x in nbs] == ref
        for n in range(len(ref)):
            for ir in (0, 1):
                for pr in (False, True):
                    assert t.calc_indices(b, n, ir, pr) == m.calc_indices(b, n, ir, pr), (b, n, ir, pr)


@pytest.mark.parametrize("ndim,nx,ng,nrb", [
    (3, (8, 8, 8), 4, (2, 2, 2)),
    (3, (16, 8, 4), 2, (4, 4, 4)),
    (2, (16, 16), 2, (8, 8)),
    (3, (32, 32, 32), 4, (4, 4, 4)),
    (1, (8,), 2, (4,)),
])
def test_uniform_topology_matches_oracle(ndim, nx, ng, nrb):
    m = oracle.Mesh(ndim, nx, ng, nrb)
    t = host.Topology(overrides=deck_overrides(ndim, nx, ng, nrb))
    compare_topology(t, m)


@pytest.mark.parametrize("nrb,refine", [(2, {(0, 0, 0)}), (4, {(1, 1, 1), (2, 1, 1)}),
                                        (2, {(0, 0, 0), (1, 1, 1)})])
def test_refined_topology_matches_oracle(nrb, refine):
    nx, ng = (8, 8, 8), 2
    leaves = H.refined_leaves(nrb, refine)
    m = oracle.Mesh(3, nx, ng, (nrb,) * 3, leaves=leaves)
    t = host.Topology(overrides=deck_overrides(3, nx, ng, (nrb,) * 3, "static"), leaves=leaves)
    assert t.info()["multilevel"] == 1
    compare_topology(t, m)


def test_rank_assignment_follows_reference():
    # mesh-amr_loadbalance.cpp:362-386: contiguous ranges, filled from the last rank down,
    # so with 27 blocks on 4 ranks rank 0 holds the short share
    ov = deck_overrides(3, (8, 8, 8), 2, (4, 4, 4))
    t = host.Topology(overrides=ov, rank=0, nranks=8)
    r = t.ranklist()
    assert np.array_equal(r, np.repeat(np.arange(8), 8))
    # 512^3 / 32^3 on 8 GPUs: one octant (512 consecutive Morton ids) per GPU
    ov = deck_overrides(3, (4, 4, 4), 2, (16, 16, 16))
    t = host.Topology(overrides=ov, rank=3, nranks=8)
    r = t.ranklist()
    assert np.array_equal(np.bincount(r), np.full(8, 512))
    first = t.info()["first_gid"]
    assert first == 3 * 512
    locs = np.array([t.block(b)["loc"] for b in range(512)])
    # rank 3 owns the octant x high, y high, z low (Morton: x lowest bit)
    assert locs[:, 1].min() >= 8 and locs[:, 2].min() >= 8 and locs[:, 3].max() < 8
    # uneven split: remainder goes to the high ranks
    leaves = H.refined_leaves(2, {(0, 0, 0)})
    t = host.Topology(overrides=deck_overrides(3, (8, 8, 8), 2, (2, 2, 2), "static"), leaves=leaves,
                      nranks=4)
    counts = np.bincount(t.ranklist(), minlength=4)
    assert counts.sum() == 15 and list(counts) == [3, 4, 4, 4]


def plans_consistent(ov, nranks, ncomp, leaves=None):
    """every (sender rank -> receiver rank) slab segment must have identical channel order,
    offsets and sizes on both sides, and cover all inter-rank channels exactly once"""
    tops = [host.Topology(overrides=ov, rank=r, nranks=nranks, leaves=leaves)
            for r in range(nranks)]
    sends, recvs = {}, {}
    for r, t in enumerate(tops):
        rows, seg = t.plan(ncomp, "send")
        for row in rows:
            sends.setdefault((r, int(row[6])), []).append((tuple(row[:4]), row[4] - seg[row[6]], row[5]))
        rows, seg = t.plan(ncomp, "recv")
        for row in rows:
            recvs.setdefault((int(row[6]), r), []).append((tuple(row[:4]), row[4] - seg[row[6]], row[5]))
    assert sends.keys() == recvs.keys()
    # peer push: a sender addresses the receiver's slab as "the receiver's segment
    # offset for me + the channel's offset inside my segment for it" — it must be where the
    # receiver's own plan unpacks that channel from
    recv_abs = {}
    segs = []
    for r, t in enumerate(tops):
        rows, seg = t.plan(ncomp, "recv")
        segs.append(seg)
        for row in rows:
            recv_abs[(r,) + tuple(int(x) for x in row[:4])] = int(row[4])
    for r, t in enumerate(tops):
        rows, seg = t.plan(ncomp, "send")
        for row in rows:
            peer = int(row[6])
            dst = int(segs[peer][r]) + int(row[4] - seg[peer])
            assert dst == recv_abs[(peer,) + tuple(int(x) for x in row[:4])]
    total = 0
    for key in sends:
        assert sends[key] == recvs[key], key
        total += sum(x[2] for x in sends[key])
    return total, tops


def test_slab_layouts_agree_between_ranks():
    ov = deck_overrides(3, (8, 8, 8), 4, (4, 4, 4))
    total, tops = plans_consistent(ov, 4, 3)
    # uniform mesh: every ghost cell is filled exactly once, locally or through a slab
    local = sum(int(t.plan(3, "local")[0][:, 5].sum()) for t in tops)
    ghosts = 64 * 3 * (16 ** 3 - 8 ** 3)
    assert total + local == ghosts
    # multilevel
    leaves = H.refined_leaves(2, {(0, 0, 0)})
    plans_consistent(deck_overrides(3, (8, 8, 8), 2, (2, 2, 2), "static"), 3, 2, leaves)


def test_virtual_ranks_split_local_and_slab_channels():
    ov = deck_overrides(3, (8, 8, 8), 4, (2, 2, 2))
    t1 = host.Topology(overrides=ov)
    ov2 = dict(ov)
    ov2["pb2/virtual_ranks"] = 2
    t2 = host.Topology(overrides=ov2)
    l1 = t1.plan(1, "local")[0]
    l2, s2, r2 = (t2.plan(1, k)[0] for k in ("local", "send", "recv"))
    assert len(l1) == 8 * 26 and len(t1.plan(1, "send")[0]) == 0
    assert len(l2) + len(r2) == len(l1) and len(s2) == len(r2) > 0
    # send and receive slabs of the virtual ranks share one layout
    assert np.array_equal(s2[:, :6], r2[:, :6])


def test_static_refinement_from_deck():
    deck = host.BURGERS_DECK + """
<parthenon/static_refinement0>
x1min = -0.5
x1max = -0.25
x2min = -0.5
x2max = -0.25
x3min = -0.5
x3max = -0.25
level = 1
"""
    ov = deck_overrides(3, (8, 8, 8), 2, (4, 4, 4), "static")
    t = host.Topology(deck=deck, overrides=ov)
    info = t.info()
    assert info["multilevel"] == 1 and info["nbtotal"] == 64 - 1 + 8
    levels = [t.block(b)["loc"][0] for b in range(info["nbtotal"])]
    assert levels.count(3) == 8 and levels.count(2) == 63


def test_elongated_mesh_partitions_into_cubes():
    """512x256x256-style root grids (bench.py's weak-scaling meshes): Morton order in the
    enclosing cube gives each of 2 / 4 ranks one compact cubic brick of blocks"""
    for nrb, nranks in (((4, 2, 2), 2), ((4, 4, 2), 4)):
        ov = deck_overrides(3, (4, 4, 4), 2, nrb)
        for rank in range(nranks):
            t = host.Topology(overrides=ov, rank=rank, nranks=nranks)
            info = t.info()
            assert info["nbtotal"] == nrb[0] * nrb[1] * nrb[2] and info["nblocks"] == 8
            locs = np.array([t.block(b)["loc"][1:] for b in range(8)])
            assert (locs.max(axis=0) - locs.min(axis=0) == 1).all()  # a 2x2x2 brick
            for b in range(8):
                assert len(t.neighbors(b)) == 26
        plans_consistent(ov, nranks, 2)
    # coordinates of an elongated mesh still tile the domain exactly
    t = host.Topology(overrides=deck_overrides(3, (4, 4, 4), 2, (4, 2, 2)))
    xs = sorted({(t.block(b)["xmin"][0], t.block(b)["xmax"][0]) for b in range(16)})
    assert xs[0][0] == -0.5 and xs[-1][1] == 0.5
    assert all(xs[i][1] == xs[i + 1][0] for i in range(3))


@pytest.mark.parametrize("ndim,nb,nrb,numlevel,dc,ncyc", [(2, 8, 4, 3, 3, 40), (3, 8, 4, 3, 3, 14)])
def test_adaptive_tree_update_follows_oracle(ndim, nb, nrb, numlevel, dc, ncyc):
    """the host half of a remesh without a device: the raw refinement tags of the CPU oracle's
    adaptive advection run (pinned to the reference) are fed, cycle by cycle, to the host
    library's SetRefinement / UpdateMeshBlockTree / block-list rebuild; block lists must agree
    after every regrid (level limits, derefinement counters that survive on kept blocks,
    sibling-complete derefinement, proper nesting)"""
    A = oracle.AmrAdvection(ndim, (nb,) * ndim, 2, (nrb,) * ndim, numlevel, derefine_count=dc)
    ov = deck_overrides(ndim, (nb,) * 3, 2, (nrb,) * 3, refinement="adaptive")
    ov.update({"parthenon/mesh/numlevel": numlevel, "parthenon/mesh/derefine_count": dc})

    def same():
        n = t.info()["nblocks"]
        return n == A.nblocks and np.array_equal(
            np.array([t.block(b)["loc"] for b in range(n)]), A.block_locs)

    # both start from the oracle's mesh and counters after its initial refinement loop
    A.init()
    t = host.Topology(overrides=ov, leaves=A.block_locs)
    t.derefine_counts = A.deref_counts
    assert same()
    changes = 0
    for c in range(ncyc):
        A.step()
        tags = A.tags
        ch_o = A.regrid()
        ch_h = t.regrid(tags)
        assert ch_o == ch_h, c
        assert same(), c
        assert np.array_equal(t.derefine_counts, A.deref_counts), c
        changes += ch_o
    assert changes > 2


def apply_plan_rows(rows, src, dst, first_gid=0):
    """what the copy kernel does with the host's channel pieces: dst box <- src box"""
    for r in rows:
        sg, rg, c0, nc = int(r[0]) - first_gid, int(r[1]) - first_gid, int(r[4]), int(r[5])
        (si, sj, sk), (ri, rj, rk), (ni, nj, nk) = r[6:9], r[9:12], r[12:15]
        dst[rg, c0:c0 + nc, rk:rk + nk, rj:rj + nj, ri:ri + ni] = \
            src[sg, c0:c0 + nc, sk:sk + nk, sj:sj + nj, si:si + ni]


@pytest.mark.parametrize("name,ndim,nx,nb,ng", H.TECOMM)
def test_non_cell_centred_plan_reproduces_reference(name, ndim, nx, nb, ng):
    """face / edge / node exchange, host half without a device: the channel pieces the host
    library derives (element boxes, ownership masks resolved into sub-boxes) are applied with
    numpy to the reference problem generator's state and must give the reference's dump bit
    for bit.  Pieces read only entries their sender owns and never overlap on the receiver, so
    the one-launch fused copy needs no ordering."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    nrb = nx // nb
    ov = deck_overrides(ndim, (nb,) * 3, ng, (nrb,) * 3)
    t = host.Topology(overrides=ov)
    m = oracle.Mesh(ndim, (nb,) * ndim, ng, (nrb,) * ndim)
    for kind, key, ncomp in H.TECOMM_FIELDS:
        nel = 3 if kind < 3 else 1
        nk, nj, ni = m.te_extents(kind)
        ref = g[key]
        U0 = H.tecomm_initial(m.nblocks, nel, ncomp, nk, nj, ni).reshape(ref.shape)
        rows = t.plan_boxes(ncomp, kind, "local")
        U = U0.copy()
        apply_plan_rows(rows, U0, U)
        assert np.array_equal(U, ref), (name, key)
        # no entry is written twice, and nothing that is read is also written (race freedom of
        # the fused copy)
        written = np.zeros(ref.shape, dtype=np.int32)
        read = np.zeros(ref.shape, dtype=bool)
        for r in rows:
            sg, rg, c0, nc = int(r[0]), int(r[1]), int(r[4]), int(r[5])
            (si, sj, sk), (ri, rj, rk), (bi, bj, bk) = r[6:9], r[9:12], r[12:15]
            written[rg, c0:c0 + nc, rk:rk + bk, rj:rj + bj, ri:ri + bi] += 1
            read[sg, c0:c0 + nc, sk:sk + bk, sj:sj + bj, si:si + bi] = True
        assert written.max() == 1
        assert not (read & (written > 0)).any()
        # everything that is written changes (the codes are block dependent): exactly the
        # ghosts and the shared elements a block does not own
        assert np.array_equal(written > 0, U != U0)


@pytest.mark.parametrize("name,ndim,nx,nb,ng", [H.TECOMM[0], H.TECOMM[2], H.TECOMM[3]])
def test_non_cell_centred_plan_two_ranks(name, ndim, nx, nb, ng):
    """the same across two devices, still without one: each rank derives its send and receive
    pieces from topology alone (the receiver computes the SENDER's ownership mask from the tree
    every rank holds); packing rank A's send pieces into a slab and unpacking them with rank B's
    receive pieces, plus both ranks' local pieces, must again give the reference's dump"""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    nrb = nx // nb
    ov = deck_overrides(ndim, (nb,) * 3, ng, (nrb,) * 3)
    topo = [host.Topology(overrides=ov, rank=r, nranks=2) for r in range(2)]
    m = oracle.Mesh(ndim, (nb,) * ndim, ng, (nrb,) * ndim)
    for kind, key, ncomp in H.TECOMM_FIELDS:
        nel = 3 if kind < 3 else 1
        nk, nj, ni = m.te_extents(kind)
        ref = g[key]
        U0 = H.tecomm_initial(m.nblocks, nel, ncomp, nk, nj, ni).reshape(ref.shape)
        U = U0.copy()
        slabs = {}
        for r in range(2):
            apply_plan_rows(topo[r].plan_boxes(ncomp, kind, "local"), U0, U)
            send = topo[r].plan_boxes(ncomp, kind, "send")
            assert len(send) > 0
            total = int(max(s[15] + s[5] * s[12] * s[13] * s[14] for s in send)) + 1
            slab = np.full(total, np.nan)
            for s in send:
                sg, c0, nc = int(s[0]), int(s[4]), int(s[5])
                (si, sj, sk), (bi, bj, bk) = s[6:9], s[12:15]
                box = U0[sg, c0:c0 + nc, sk:sk + bk, sj:sj + bj, si:si + bi]
                slab[int(s[15]):int(s[15]) + box.size] = box.ravel()
            slabs[r] = slab
        for r in range(2):
            recv = topo[r].plan_boxes(ncomp, kind, "recv")
            send = topo[1 - r].plan_boxes(ncomp, kind, "send")
            # the two sides list the same pieces in the same order at the same offsets
            assert np.array_equal(recv[:, [0, 1, 2, 3, 4, 5, 12, 13, 14, 15]],
                                  send[:, [0, 1, 2, 3, 4, 5, 12, 13, 14, 15]])
            for q in recv:
                rg, c0, nc = int(q[1]), int(q[4]), int(q[5])
                (ri, rj, rk), (bi, bj, bk) = q[9:12], q[12:15]
                n = nc * bi * bj * bk
                U[rg, c0:c0 + nc, rk:rk + bk, rj:rj + bj, ri:ri + bi] = \
                    slabs[1 - r][int(q[15]):int(q[15]) + n].reshape(nc, bk, bj, bi)
        assert np.array_equal(U, ref), (name, key)


@pytest.mark.parametrize("name,ndim,nx,nb,ng", H.TECOMM_MULTILEVEL)
def test_non_cell_centred_multilevel_plan_matches_oracle_regions(name, ndim, nx, nb, ng):
    """refined meshes, host half without a device: for every block the entries its local channel
    pieces write — into the fine array or, from a coarser sender, into the coarse buffer — are
    exactly the entries the oracle's masked unpack (pinned to the reference) writes; pieces never
    overlap, and what they read on the sender is exactly the sender-side box entries the
    receiver's mask lets through"""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    full = lambda n: (n,) * ndim + (1,) * (3 - ndim)
    leaves, nrb = H.leaves_from_bounds(g["bounds"], full(nx), full(nb))
    ov = deck_overrides(ndim, (nb,) * 3, ng, nrb, refinement="static")
    t = host.Topology(overrides=ov, leaves=leaves)
    m = oracle.Mesh(ndim, (nb,) * ndim, ng, tuple(nrb[:ndim]), leaves=leaves)
    for kind in (1, 2, 3):
        nel = 3 if kind < 3 else 1
        fine = (m.nblocks, nel) + m.te_extents(kind)
        coarse = (m.nblocks, nel) + tuple(n + (1 if n > 1 else 0) for n in m.cdims)
        want = {False: np.zeros(fine, dtype=np.int32), True: np.zeros(coarse, dtype=np.int32)}
        for b in range(m.nblocks):
            lev = m.block_loc(b)[0]
            for n, nbr in enumerate(m.neighbors(b)):
                for el in range(nel):
                    s, e, mask = m.calc_indices_te_general(b, n, kind, el, 1)
                    to_coarse = nbr[1] < lev  # (gid, level, ox1, ox2, ox3)
                    for k in range(s[2], e[2] + 1):
                        kk = (k == e[2]) - (k == s[2])
                        for j in range(s[1], e[1] + 1):
                            jj = (j == e[1]) - (j == s[1])
                            for i in range(s[0], e[0] + 1):
                                ii = (i == e[0]) - (i == s[0])
                                if mask[kk + 1, jj + 1, ii + 1]:
                                    want[to_coarse][b, el, k, j, i] = 1
        got = {False: np.zeros(fine, dtype=np.int32), True: np.zeros(coarse, dtype=np.int32)}
        for r in t.plan_boxes(1, kind, "local"):
            rg, el = int(r[1]), int(r[4])
            (ri, rj, rk), (bi, bj, bk) = r[9:12], r[12:15]
            got[bool(int(r[17]) & 2)][rg, el, rk:rk + bk, rj:rj + bj, ri:ri + bi] += 1
        for to_coarse in (False, True):
            assert got[to_coarse].max() <= 1, (kind, to_coarse)
            assert np.array_equal(got[to_coarse] > 0, want[to_coarse] > 0), (kind, to_coarse)
        assert got[True].sum() > 0 and got[False].sum() > 0


TEFLUX = [("teflux_s16_b8_l2_3d", 3, 16, 8, 2), ("teflux_s32_b8_l3_2d", 2, 32, 8, 2),
          ("teflux_s16_b4_g4_l3_3d_sparse", 3, 16, 4, 4)]


@pytest.mark.parametrize("name,ndim,nx,nb,ng", TEFLUX)
def test_edge_flux_plan_reproduces_oracle(name, ndim, nx, nb, ng):
    """flux correction of a face field, host half without a device: the restrict boxes the host
    library derives are exactly the coarse-buffer entries the oracle (pinned to three reference
    dumps) restricts, and its delivery pieces — sender's coarse buffer -> receiver's flux array,
    block-edge messages first, face messages second — applied with numpy to the oracle's coarse
    buffers give the oracle's corrected flux field bit for bit"""
    import ctypes as C
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    full = lambda n: (n,) * ndim + (1,) * (3 - ndim)
    leaves, nrb = H.leaves_from_bounds(g["bounds"], full(nx), full(nb))
    ov = deck_overrides(ndim, (nb,) * 3, ng, nrb, refinement="static")
    t = host.Topology(overrides=ov, leaves=leaves)
    m = oracle.Mesh(ndim, (nb,) * ndim, ng, tuple(nrb[:ndim]), leaves=leaves)
    nk, nj, ni = m.te_extents(2)
    gid = np.arange(m.nblocks).reshape(-1, 1, 1, 1, 1, 1)
    e = np.arange(3).reshape(1, -1, 1, 1, 1, 1)
    F0 = ((gid + 1) * 1.0e6 + e * 1.0e5 +
          np.arange(nk * nj * ni).reshape(1, 1, 1, nk, nj, ni)).astype(np.float64)
    F = F0.copy()
    cd = tuple(n + (1 if n > 1 else 0) for n in m.cdims)
    Fc = np.zeros(F.shape[:3] + cd)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    assert oracle.lib().orc_flux_correct_edge(m.h, dp(F), dp(Fc), 1, None) > 0
    # restrict boxes == what the oracle restricted (the codes are positive, Fc started at zero)
    rs = t.edge_flux_plan("restrict")
    want = np.zeros(Fc.shape, dtype=bool)
    for r in rs:
        b, el = int(r[0]), int(r[2])
        (si, sj, sk), (bi, bj, bk) = r[4:7], r[10:13]
        want[b, el, 0, sk:sk + bk, sj:sj + bj, si:si + bi] = True
    assert np.array_equal(want, Fc != 0)
    # deliveries in pass order
    ds = t.edge_flux_plan("deliver")
    assert len(ds) > 0 and set(np.unique(ds[:, 3])) <= {0, 1}
    G = F0.copy()
    written = np.zeros(F.shape, dtype=np.int32)
    for p in (0, 1):
        for r in ds[ds[:, 3] == p]:
            sg, rg, el = int(r[0]), int(r[1]), int(r[2])
            (si, sj, sk), (ri, rj, rk), (bi, bj, bk) = r[4:7], r[7:10], r[10:13]
            src = Fc[sg, el, 0, sk:sk + bk, sj:sj + bj, si:si + bi]
            assert np.all(src != 0)  # only restricted entries travel
            G[rg, el, 0, rk:rk + bk, rj:rj + bj, ri:ri + bi] = src
            written[rg, el, 0, rk:rk + bk, rj:rj + bj, ri:ri + bi] += 1
    assert np.array_equal(G, F)
    # within one pass nothing is written twice, so each pass is one race-free launch
    for p in (0, 1):
        w = np.zeros(F.shape, dtype=np.int32)
        for r in ds[ds[:, 3] == p]:
            rg, el = int(r[1]), int(r[2])
            (ri, rj, rk), (bi, bj, bk) = r[7:10], r[10:13]
            w[rg, el, 0, rk:rk + bk, rj:rj + bj, ri:ri + bi] += 1
        assert w.max() <= 1


@pytest.mark.parametrize("nranks", [2, 3])
@pytest.mark.parametrize("name,ndim,nx,nb,ng", TEFLUX)
def test_edge_flux_plan_across_ranks(name, ndim, nx, nb, ng, nranks):
    """the same with the blocks spread over several devices, still without one: every rank
    derives what its fine blocks restrict and send and what its coarse blocks receive from the
    topology alone; a peer's send list and this rank's receive list name the same pieces in the
    same order at the same slab offsets, and same-device copies plus slab traffic — block-edge
    messages before face messages — give the oracle's corrected flux field on every rank"""
    import ctypes as C
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    full = lambda n: (n,) * ndim + (1,) * (3 - ndim)
    leaves, nrb = H.leaves_from_bounds(g["bounds"], full(nx), full(nb))
    ov = deck_overrides(ndim, (nb,) * 3, ng, nrb, refinement="static")
    topo = [host.Topology(overrides=ov, leaves=leaves, rank=r, nranks=nranks) for r in range(nranks)]
    m = oracle.Mesh(ndim, (nb,) * ndim, ng, tuple(nrb[:ndim]), leaves=leaves)
    nk, nj, ni = m.te_extents(2)
    gid = np.arange(m.nblocks).reshape(-1, 1, 1, 1, 1, 1)
    e = np.arange(3).reshape(1, -1, 1, 1, 1, 1)
    F0 = ((gid + 1) * 1.0e6 + e * 1.0e5 +
          np.arange(nk * nj * ni).reshape(1, 1, 1, nk, nj, ni)).astype(np.float64)
    F = F0.copy()
    cd = tuple(n + (1 if n > 1 else 0) for n in m.cdims)
    Fc = np.zeros(F.shape[:3] + cd)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    assert oracle.lib().orc_flux_correct_edge(m.h, dp(F), dp(Fc), 1, None) > 0
    ranklist = topo[0].ranklist()
    # restrictions: local ones listed by the receivers + those of senders to other devices
    want = np.zeros(Fc.shape, dtype=bool)
    for r in range(nranks):
        for kind in ("restrict", "send_restrict"):
            for row in topo[r].edge_flux_plan(kind):
                b, el = int(row[0]), int(row[2])
                assert ranklist[b] == r
                (si, sj, sk), (bi, bj, bk) = row[4:7], row[10:13]
                want[b, el, 0, sk:sk + bk, sj:sj + bj, si:si + bi] = True
    assert np.array_equal(want, Fc != 0)
    # every rank packs its send pieces into one slab per peer
    slabs = {}
    nsend = 0
    for r in range(nranks):
        send = topo[r].edge_flux_plan("send")
        nsend += len(send)
        for peer in range(nranks):
            rows = send[send[:, 13] == peer]
            total = int(max([s[14] + s[10] * s[11] * s[12] for s in send] + [0])) + 2
            slab = np.full(total, np.nan)
            for s in rows:
                sg, el = int(s[0]), int(s[2])
                assert ranklist[sg] == r and ranklist[int(s[1])] == peer
                (si, sj, sk), (bi, bj, bk) = s[4:7], s[10:13]
                box = Fc[sg, el, 0, sk:sk + bk, sj:sj + bj, si:si + bi]
                assert np.all(box != 0)
                slab[int(s[14]):int(s[14]) + box.size] = box.ravel()
            slabs[(r, peer)] = (slab, rows)
    assert nsend > 0
    G = F0.copy()
    for r in range(nranks):
        loc, recv = topo[r].edge_flux_plan("deliver"), topo[r].edge_flux_plan("recv")
        # a rank's segments start where the peer's send slab puts them: offsets are relative to
        # the whole slab of the OTHER side, so compare them per peer after removing the base
        for peer in range(nranks):
            mine = recv[recv[:, 13] == peer]
            theirs = slabs[(peer, r)][1]
            assert len(mine) == len(theirs)
            if len(mine) == 0:
                continue
            cols = [0, 1, 2, 3, 10, 11, 12, 15]
            assert np.array_equal(mine[:, cols], theirs[:, cols])
            assert np.array_equal(mine[:, 14] - mine[0, 14], theirs[:, 14] - theirs[0, 14])
        for p in (0, 1):
            for row in loc[loc[:, 3] == p]:
                sg, rg, el = int(row[0]), int(row[1]), int(row[2])
                assert ranklist[sg] == r and ranklist[rg] == r
                (si, sj, sk), (ri, rj, rk), (bi, bj, bk) = row[4:7], row[7:10], row[10:13]
                G[rg, el, 0, rk:rk + bk, rj:rj + bj, ri:ri + bi] = \
                    Fc[sg, el, 0, sk:sk + bk, sj:sj + bj, si:si + bi]
            for peer in range(nranks):
                mine = recv[(recv[:, 13] == peer) & (recv[:, 3] == p)]
                slab, theirs = slabs[(peer, r)]
                if len(mine) == 0:
                    continue
                allmine = recv[recv[:, 13] == peer]
                shift = int(theirs[0, 14]) - int(allmine[0, 14])
                for row in mine:
                    rg, el = int(row[1]), int(row[2])
                    assert ranklist[rg] == r
                    (ri, rj, rk), (bi, bj, bk) = row[7:10], row[10:13]
                    n = bi * bj * bk
                    o = int(row[14]) + shift
                    G[rg, el, 0, rk:rk + bk, rj:rj + bj, ri:ri + bi] = \
                        slab[o:o + n].reshape(bk, bj, bi)
    assert not np.isnan(G).any()
    assert np.array_equal(G, F)


def test_plan_is_the_same_on_one_and_on_many_host_threads():
    """the exchange plan of a mesh of more than 256 blocks is built on all host threads (channels
    per block, lists joined by prefix sums): channel order, boxes and slab offsets must equal the
    single-threaded build, for same-device channels and for the slabs of a 2-rank partition"""
    import ctypes as C
    try:
        gomp = C.CDLL("libgomp.so.1")
    except OSError:
        pytest.skip("no libgomp")
    leaves = H.refined_leaves(8, {(3, 3, 3), (4, 4, 4), (1, 6, 2)})
    ov = deck_overrides(3, (8, 8, 8), 2, (8, 8, 8), "static")
    got = {}
    for nt in (1, 8):
        gomp.omp_set_num_threads(nt)
        rows = []
        for rank, nranks in ((0, 1), (0, 2), (1, 2)):
            t = host.Topology(overrides=ov, leaves=leaves, rank=rank, nranks=nranks)
            assert t.info()["nbtotal"] > 512
            rows += [t.plan_boxes(3, 0, kind) for kind in ("local", "send", "recv")]
        got[nt] = rows
    gomp.omp_set_num_threads(max(1, len(os.sched_getaffinity(0))))
    assert sum(len(r) for r in got[1]) > 10000
    for a, b in zip(got[1], got[8]):
        assert np.array_equal(a, b)
