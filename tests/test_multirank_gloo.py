"""world_size-2 test of the N > 1 host logic on CPU (gloo): every process builds ITS OWN
partition and exchange plan with the C++ host library (no device), packs the per-peer slabs the
plan prescribes from a seeded field, ships them with torch.distributed send/recv over gloo,
unpacks what it receives — and must end up with the ghosts the single-process CPU oracle
computes for the whole mesh.  This is the path NCCL carries on the GPUs (pb2_comm_exchange):
contiguous Morton gid ranges per rank, one slab segment per peer, offsets derived on both
sides from topology alone (no handshake)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from parthenon_b200 import host
from tests.test_host_topology import deck_overrides

NX, NG, NRB, NCOMP, WORLD = (8, 6, 4), 2, (4, 4, 4), 2, 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _box(m, gid, other, off, ir):
    """index box of the channel between `gid` and its neighbour `other` at offsets `off`"""
    for n, nb in enumerate(m.neighbors(gid)):
        if nb[0] == other and tuple(nb[2:]) == tuple(off):
            s, e = m.calc_indices(gid, n, ir)
            return tuple(slice(s[d], e[d] + 1) for d in (2, 1, 0))
    raise AssertionError("channel without a neighbour entry")


def _offsets(idx):
    return (idx % 3 - 1, (idx // 3) % 3 - 1, idx // 9 - 1)


def _worker(rank, port, result):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        m = oracle.Mesh(3, NX, NG, NRB)
        rng = np.random.default_rng(3)
        U = rng.standard_normal((m.nblocks, NCOMP) + m.dims)
        Uref = U.copy()
        m.exchange(Uref)

        t = host.Topology(overrides=deck_overrides(3, NX, NG, NRB), rank=rank, nranks=WORLD)
        info = t.info()
        lo, hi = info["first_gid"], info["first_gid"] + info["nblocks"]
        mine = U.copy()
        mine[:lo] = np.nan  # this process only holds its own blocks
        mine[hi:] = np.nan

        send_rows, send_seg = t.plan(NCOMP, "send")
        recv_rows, recv_seg = t.plan(NCOMP, "recv")
        send = np.zeros(int(send_seg[WORLD]))
        for sg, rg, _var, oi, off, n, _peer in send_rows:
            assert lo <= sg < hi
            # offset_index is the sender's view of the receiver
            box = _box(m, int(sg), int(rg), _offsets(int(oi)), 0)
            send[off:off + n] = mine[(int(sg), slice(None)) + box].ravel()
        recv = np.zeros(int(recv_seg[WORLD]))
        reqs = []
        for p in range(WORLD):
            if p == rank:
                continue
            a, b = int(send_seg[p]), int(send_seg[p + 1])
            if b > a:
                reqs.append(dist.isend(torch.from_numpy(send[a:b].copy()), p))
        bufs = {}
        for p in range(WORLD):
            if p == rank:
                continue
            a, b = int(recv_seg[p]), int(recv_seg[p + 1])
            if b > a:
                bufs[p] = torch.empty(b - a, dtype=torch.float64)
                dist.recv(bufs[p], p)
                recv[a:b] = bufs[p].numpy()
        for r in reqs:
            r.wait()
        for sg, rg, _var, oi, off, n, _peer in recv_rows:
            assert lo <= rg < hi and not (lo <= sg < hi)
            ro = tuple(-x for x in _offsets(int(oi)))  # the receiver sees the mirror offset
            box = _box(m, int(rg), int(sg), ro, 1)
            tgt = mine[(int(rg), slice(None)) + box]
            tgt[...] = recv[off:off + n].reshape(tgt.shape)
        for sg, rg, _var, oi, _off, n, _peer in t.plan(NCOMP, "local")[0]:
            ro = tuple(-x for x in _offsets(int(oi)))
            src = mine[(int(sg), slice(None)) + _box(m, int(sg), int(rg), _offsets(int(oi)), 0)]
            tgt = mine[(int(rg), slice(None)) + _box(m, int(rg), int(sg), ro, 1)]
            tgt[...] = src
        ok = np.array_equal(mine[lo:hi], Uref[lo:hi])
        moved = torch.tensor([float(send.size)], dtype=torch.float64)
        dist.all_reduce(moved)
        result[rank] = (bool(ok), hi - lo, float(moved.item()))
    finally:
        dist.destroy_process_group()


def test_two_rank_slab_exchange_over_gloo():
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        result = mgr.dict()
        mp.spawn(_worker, args=(_free_port(), result), nprocs=WORLD, join=True)
        res = dict(result)
    assert set(res) == {0, 1}
    assert all(r[0] for r in res.values()), res
    assert sum(r[1] for r in res.values()) == int(np.prod(NRB))
    assert res[0][2] == res[1][2] > 0  # all-reduced count of Reals that crossed ranks



# ---- face / edge / node fields: channel pieces (ownership masks resolved into boxes) ----
def _te_worker(rank, port, result):
    """each process holds only its own blocks of the reference problem generator's state,
    derives its send / receive / local pieces with the host library (the receiver computes the
    SENDER's ownership from the tree), ships one slab to its peer over gloo and must end up
    with the reference's dump on its blocks"""
    from tests import helpers as H
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        name, ndim, nx, nb, ng = H.TECOMM[2]  # 4 x 4 x 4 blocks of 4^3
        g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
        nrb = nx // nb
        t = host.Topology(overrides=deck_overrides(ndim, (nb,) * 3, ng, (nrb,) * 3), rank=rank,
                          nranks=WORLD)
        info = t.info()
        lo, hi = info["first_gid"], info["first_gid"] + info["nblocks"]
        peer = 1 - rank
        ok, moved = True, 0
        size = lambda r: int(r[5] * r[12] * r[13] * r[14])
        for kind, key, ncomp in H.TECOMM_FIELDS:
            ref = g[key]
            nblocks, _, nk, nj, ni = ref.shape
            nel = 3 if kind < 3 else 1
            U = H.tecomm_initial(nblocks, nel, ncomp, nk, nj, ni).reshape(ref.shape)
            U[:lo] = np.nan  # the other rank's blocks are not here
            U[hi:] = np.nan
            U0 = U.copy()
            send_rows = t.plan_boxes(ncomp, kind, "send")
            recv_rows = t.plan_boxes(ncomp, kind, "recv")
            send = np.full(max(int(r[15]) + size(r) for r in send_rows), np.nan)
            for r in send_rows:
                sg, c0, nc = int(r[0]), int(r[4]), int(r[5])
                (si, sj, sk), (bi, bj, bk) = r[6:9], r[12:15]
                send[int(r[15]):int(r[15]) + size(r)] = \
                    U0[sg, c0:c0 + nc, sk:sk + bk, sj:sj + bj, si:si + bi].ravel()
            recv = torch.empty(max(int(r[15]) + size(r) for r in recv_rows), dtype=torch.float64)
            req = dist.isend(torch.from_numpy(send.copy()), peer)
            dist.recv(recv, peer)
            req.wait()
            recv = recv.numpy()
            for r in recv_rows:
                rg, c0, nc = int(r[1]), int(r[4]), int(r[5])
                (ri, rj, rk), (bi, bj, bk) = r[9:12], r[12:15]
                U[rg, c0:c0 + nc, rk:rk + bk, rj:rj + bj, ri:ri + bi] = \
                    recv[int(r[15]):int(r[15]) + size(r)].reshape(nc, bk, bj, bi)
            for r in t.plan_boxes(ncomp, kind, "local"):
                sg, rg, c0, nc = int(r[0]), int(r[1]), int(r[4]), int(r[5])
                (si, sj, sk), (ri, rj, rk), (bi, bj, bk) = r[6:9], r[9:12], r[12:15]
                U[rg, c0:c0 + nc, rk:rk + bk, rj:rj + bj, ri:ri + bi] = \
                    U0[sg, c0:c0 + nc, sk:sk + bk, sj:sj + bj, si:si + bi]
            ok = ok and np.array_equal(U[lo:hi], ref[lo:hi])
            moved += send.size
        result[rank] = (bool(ok), hi - lo, moved)
    finally:
        dist.destroy_process_group()


def test_two_rank_face_edge_node_exchange_over_gloo():
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        result = mgr.dict()
        mp.spawn(_te_worker, args=(_free_port(), result), nprocs=WORLD, join=True)
        res = dict(result)
    assert set(res) == {0, 1}
    assert all(r[0] for r in res.values()), res
    assert sum(r[1] for r in res.values()) == 64
    assert res[0][2] > 0 and res[1][2] > 0


# ---- flux correction of a face field: restricted edge fluxes cross ranks in one slab ----
def _edge_flux_worker(rank, port, result):
    """each process holds only the coarse buffers (restricted edge fluxes, from the oracle) and
    the flux arrays of ITS blocks of a statically refined mesh, derives with the host library
    what it sends and receives, ships one slab to its peer over gloo and applies same-device
    copies and the received pieces in the two passes — block-edge messages, then face messages.
    Its blocks must hold the oracle's corrected flux field (pinned to the reference's dump)"""
    import ctypes as C
    from tests import helpers as H
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        name, ndim, nx, nb, ng = "teflux_s16_b8_l2_3d", 3, 16, 8, 2
        g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
        leaves, nrb = H.leaves_from_bounds(g["bounds"], (nx,) * 3, (nb,) * 3)
        ov = deck_overrides(ndim, (nb,) * 3, ng, nrb, refinement="static")
        t = host.Topology(overrides=ov, leaves=leaves, rank=rank, nranks=WORLD)
        info = t.info()
        lo, hi = info["first_gid"], info["first_gid"] + info["nblocks"]
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
        # this process: its own blocks only
        G = F0.copy()
        for a in (G, Fc):
            a[:lo] = np.nan
            a[hi:] = np.nan
        peer = 1 - rank
        size = lambda r: int(r[10] * r[11] * r[12])
        send_rows, recv_rows = t.edge_flux_plan("send"), t.edge_flux_plan("recv")
        assert len(send_rows) + len(recv_rows) > 0
        assert all(int(r[13]) == peer for r in list(send_rows) + list(recv_rows))
        base_s = min([int(r[14]) for r in send_rows] + [0])
        base_r = min([int(r[14]) for r in recv_rows] + [0])
        send = np.full(max([int(r[14]) - base_s + size(r) for r in send_rows] + [1]), np.nan)
        for r in send_rows:
            sg, el = int(r[0]), int(r[2])
            assert lo <= sg < hi
            (si, sj, sk), (bi, bj, bk) = r[4:7], r[10:13]
            o = int(r[14]) - base_s
            send[o:o + size(r)] = Fc[sg, el, 0, sk:sk + bk, sj:sj + bj, si:si + bi].ravel()
        recv = torch.full((max([int(r[14]) - base_r + size(r) for r in recv_rows] + [1]),), np.nan,
                          dtype=torch.float64)
        req = dist.isend(torch.from_numpy(send.copy()), peer)
        dist.recv(recv, peer)
        req.wait()
        recv = recv.numpy()[:max([int(r[14]) - base_r + size(r) for r in recv_rows] + [0])] \
            if len(recv_rows) else recv.numpy()[:0]
        loc = t.edge_flux_plan("deliver")
        for p in (0, 1):
            for r in loc[loc[:, 3] == p]:
                sg, rg, el = int(r[0]), int(r[1]), int(r[2])
                (si, sj, sk), (ri, rj, rk), (bi, bj, bk) = r[4:7], r[7:10], r[10:13]
                G[rg, el, 0, rk:rk + bk, rj:rj + bj, ri:ri + bi] = \
                    Fc[sg, el, 0, sk:sk + bk, sj:sj + bj, si:si + bi]
            for r in recv_rows[recv_rows[:, 3] == p] if len(recv_rows) else []:
                rg, el = int(r[1]), int(r[2])
                assert lo <= rg < hi
                (ri, rj, rk), (bi, bj, bk) = r[7:10], r[10:13]
                o = int(r[14]) - base_r
                G[rg, el, 0, rk:rk + bk, rj:rj + bj, ri:ri + bi] = \
                    recv[o:o + size(r)].reshape(bk, bj, bi)
        ok = not np.isnan(G[lo:hi]).any() and np.array_equal(G[lo:hi], F[lo:hi])
        result[rank] = (bool(ok), hi - lo, int(len(send_rows)), int(len(recv_rows)))
    finally:
        dist.destroy_process_group()


def test_two_rank_edge_flux_correction_over_gloo():
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        result = mgr.dict()
        mp.spawn(_edge_flux_worker, args=(_free_port(), result), nprocs=WORLD, join=True)
        res = dict(result)
    assert set(res) == {0, 1}
    assert all(r[0] for r in res.values()), res
    assert sum(r[1] for r in res.values()) == 15
    # what one rank sends the other receives
    assert res[0][2] == res[1][3] and res[1][2] == res[0][3] and res[0][2] + res[1][2] > 0
