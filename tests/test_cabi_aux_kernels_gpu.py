"""GPU parity tests of the auxiliary C-ABI entry points (refinement tagging reductions, sparse
block masks, donor-cell fluxes, the two halves of a sparse exchange) against numpy restatements
of the reference lines they replace and against the CPU oracle.  Bit-exact: max / min / compare
are order-independent and the arithmetic kernels are compiled with -fmad=false."""
import ctypes as C

import numpy as np
import pytest
import torch

import oracle
from parthenon_b200 import capi
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def geom_for(m, ncomp, dx_t):
    return H.make_geom(m, ncomp, dx_t)


@pytest.mark.parametrize("ndim,nx,ng", [(3, (8, 6, 4), 2), (2, (16, 8), 3)])
def test_block_minmax_derivative_quiet(ndim, nx, ng):
    nrb = (2,) * ndim
    m = oracle.Mesh(ndim, nx, ng, nrb)
    ncomp = 3
    rng = np.random.default_rng(4)
    U = rng.standard_normal((m.nblocks, ncomp) + m.dims)
    U[1] *= 1e-9  # a quiet block
    U[2] = 0.0
    dx, _ = H.block_dx(m)
    dxd = torch.from_numpy(dx).to(DEV)
    Ud = torch.from_numpy(U).to(DEV)
    g = geom_for(m, ncomp, dxd)
    L = capi.lib()
    # Kokkos::MinMax over the entire extents (advection_package.cpp:252-263)
    mm = torch.zeros(2 * m.nblocks, dtype=torch.float64, device=DEV)
    capi.check(L.pb2_block_minmax(C.byref(g), Ud.data_ptr(), None, mm.data_ptr(), None))
    ref = np.stack([U.reshape(m.nblocks, -1).min(1), U.reshape(m.nblocks, -1).max(1)], 1).ravel()
    assert np.array_equal(mm.cpu().numpy(), ref)
    # SparseDealloc's test (update.cpp:161-186): every |x| <= threshold
    q = torch.full((m.nblocks,), -1, dtype=torch.int32, device=DEV)
    mask = torch.ones(m.nblocks, dtype=torch.int32, device=DEV)
    mask[3] = 0
    capi.check(L.pb2_block_quiet_flags(C.byref(g), Ud.data_ptr(), 1e-6, mask.data_ptr(),
                                       q.data_ptr(), None))
    refq = (np.abs(U).reshape(m.nblocks, -1).max(1) <= 1e-6).astype(np.int32)
    refq[3] = -1  # masked out: untouched
    assert np.array_equal(q.cpu().numpy(), refq)
    # Refinement::FirstDerivative / SecondDerivative (refinement_package.cpp:92-150), interior
    sl = tuple(slice(ng, -ng) if m.dims[d] > 1 else slice(None) for d in range(3))
    for order in (1, 2):
        for comp in (0, 2):
            out = torch.zeros(m.nblocks, dtype=torch.float64, device=DEV)
            capi.check(L.pb2_block_derivative(C.byref(g), Ud.data_ptr(), comp, order,
                                              out.data_ptr(), None))
            q3 = U[:, comp]
            c = q3[(slice(None),) + sl]
            best = np.zeros(m.nblocks)
            for axis in range(3):
                if m.dims[axis] == 1:
                    continue
                hi = np.roll(q3, -1, axis=axis + 1)[(slice(None),) + sl]
                lo = np.roll(q3, 1, axis=axis + 1)[(slice(None),) + sl]
                if order == 1:
                    d = 0.5 * np.abs(hi - lo) / (np.abs(c) + 1.0e-20)
                else:
                    qavg = 0.5 * (hi + lo)
                    d = np.abs(qavg - c) / (np.abs(qavg) + (np.abs(c) + 1.0e-20))
                best = np.maximum(best, d.reshape(m.nblocks, -1).max(1))
            assert np.array_equal(out.cpu().numpy(), best), (order, comp)


def test_masked_dense_updates_and_donor_cell_fluxes():
    m = oracle.Mesh(3, (8, 8, 8), 2, (2, 2, 2))
    A = oracle.Advection(m, vec_size=2, profile="smooth_gaussian", amp=1.0, v=(1.0, -0.7, 0.4))
    A.init()
    U = A.U.copy()
    A.calculate_fluxes(U)
    Fref = [A.flux(d).copy() for d in range(3)]
    dx, _ = H.block_dx(m)
    dxd = torch.from_numpy(dx).to(DEV)
    g = geom_for(m, 2, dxd)
    L = capi.lib()
    Ud = torch.from_numpy(U).to(DEV)
    Fd = [torch.full_like(Ud, 7.0) for _ in range(3)]
    fl = (C.c_void_p * 3)(*[f.data_ptr() for f in Fd])
    v = (C.c_double * 3)(1.0, -0.7, 0.4)
    mask_h = np.array([1, 0, 1, 1, 0, 1, 1, 1], dtype=np.int32)
    mask = torch.from_numpy(mask_h).to(DEV)
    capi.check(L.pb2_advection_fluxes_blocks(C.byref(g), Ud.data_ptr(), fl, v, mask.data_ptr(), None))
    ng = 2
    for d in range(3):
        got = Fd[d].cpu().numpy()
        hi = [slice(ng, ng + 8 + (1 if a == d else 0)) for a in (2, 1, 0)]  # (k, j, i) face ranges
        sel = (slice(None), slice(None), hi[0], hi[1], hi[2])
        on = mask_h.astype(bool)
        assert np.array_equal(got[on][sel[1:]], Fref[d][on][sel[1:]]), d
        assert np.all(got[~on] == 7.0)  # unallocated blocks untouched (IsAllocated guards)
    # FluxDivergence / WeightedSumData with the same mask
    capi.check(L.pb2_advection_fluxes(C.byref(g), Ud.data_ptr(), fl, v, None))
    dudt = torch.full_like(Ud, 3.0)
    capi.check(L.pb2_flux_divergence_blocks(C.byref(g), fl, dudt.data_ptr(), mask.data_ptr(), None))
    full = torch.full_like(Ud, 3.0)
    capi.check(L.pb2_flux_divergence(C.byref(g), fl, full.data_ptr(), None))
    dn, fn = dudt.cpu().numpy(), full.cpu().numpy()
    assert np.array_equal(dn[mask_h == 1], fn[mask_h == 1]) and np.all(dn[mask_h == 0] == 3.0)
    z = torch.full_like(Ud, 5.0)
    capi.check(L.pb2_weighted_sum_blocks(C.byref(g), Ud.data_ptr(), dudt.data_ptr(), 0.5, 0.25,
                                         z.data_ptr(), mask.data_ptr(), None))
    zn = z.cpu().numpy()
    assert np.array_equal(zn[mask_h == 1], (0.5 * U + 0.25 * dn)[mask_h == 1])
    assert np.all(zn[mask_h == 0] == 5.0)


def test_sparse_exchange_halves():
    """pb2_copy_flags / pb2_copy_select on a uniform mesh: a message is null when the sender is
    unallocated or every |x| in its send box is below the threshold; allocated receivers get the
    data or the sparse default, unallocated receivers are skipped (boundary_communication.cpp:
    95-157, 273-334)"""
    m = oracle.Mesh(3, (8, 8, 8), 2, (2, 2, 2))
    rng = np.random.default_rng(9)
    U = rng.standard_normal((m.nblocks, 1) + m.dims)
    U[1] *= 1e-8                       # allocated but everywhere below the threshold
    alloc = np.array([1, 1, 0, 1, 1, 1, 0, 1], dtype=bool)  # blocks 2 and 6 not allocated
    thr, default = 1e-5, -2.5
    Ud = torch.from_numpy(U).to(DEV)
    regs = H.build_copy_table(m, Ud, 1)
    recv = H.region_boxes(m, 1)
    sends = H.region_boxes(m, 0)
    for i, r in enumerate(regs):
        b, _n, nb, _s, _e = recv[i]
        sb = nb[0]
        r.flag_slot = i
        r.status = (capi.REGION_ALLOCATED if alloc[sb] else 0) | \
                   (0 if alloc[b] else capi.REGION_DST_UNALLOCATED)
        r.threshold = thr
        r.default_value = default
    t = capi.Table(regs, "copy")
    flags = torch.zeros(len(regs), dtype=torch.int32, device=DEV)
    L = capi.lib()
    capi.check(L.pb2_copy_flags(t.h, flags.data_ptr(), None))
    torch.cuda.synchronize()
    assert np.array_equal(Ud.cpu().numpy(), U)  # the sender half writes nothing
    fl = flags.cpu().numpy()
    ref = U.copy()
    for i, (b, n, nb, s, ext) in enumerate(recv):
        sb = nb[0]
        sidx = H.match_send_region(m, b, nb)
        _sb, _sn, _snb, ss, sext = sends[sidx]
        src = U[sb, 0, ss[2]:ss[2] + sext[2], ss[1]:ss[1] + sext[1], ss[0]:ss[0] + sext[0]]
        nonnull = bool(alloc[sb] and (np.abs(src) >= thr).any())
        assert fl[i] == int(nonnull), i
        if alloc[b]:
            ref[b, 0, s[2]:s[2] + ext[2], s[1]:s[1] + ext[1], s[0]:s[0] + ext[0]] = \
                src if nonnull else default
    capi.check(L.pb2_copy_select(t.h, flags.data_ptr(), None))
    torch.cuda.synchronize()
    # ghost sources are interiors, which the exchange never writes: order does not matter
    assert np.array_equal(Ud.cpu().numpy(), ref)


def test_peer_push_primitives_with_the_device_as_its_own_peer():
    """pb2_peer_handshake / pb2_copy_signal / pb2_peer_signal / pb2_peer_wait through the C ABI,
    one rank whose only peer is itself: the ready flag is raised and seen, the copy lands, the last
    thread block raises the arrival flag and leaves the counter at zero, the wait returns; exchange
    numbers only ever grow"""
    from parthenon_b200 import capi
    L = capi.lib()
    n, ncomp = 12, 3
    rng = np.random.default_rng(3)
    src = torch.from_numpy(rng.standard_normal((ncomp, n, n, n))).to("cuda:0")
    dst = torch.zeros_like(src)
    r = capi.CopyRegion()
    r.src, r.dst = src.data_ptr(), dst.data_ptr()
    r.ss[:] = (1, 2, 3)
    r.ds[:] = (4, 0, 5)
    r.n[:] = (6, 7, 4)
    r.ncomp = ncomp
    r.src_stride_j = r.dst_stride_j = n
    r.src_stride_k = r.dst_stride_k = n * n
    r.src_stride_c = r.dst_stride_c = n * n * n
    r.flag_slot = -1
    r.status = capi.REGION_ALLOCATED
    t = capi.Table([r], "copy")
    flags = torch.zeros(2, dtype=torch.int32, device="cuda:0")      # [ready][arrival] x 1 rank
    counter = torch.zeros(1, dtype=torch.int32, device="cuda:0")
    peers = torch.zeros(1, dtype=torch.int32, device="cuda:0")
    peer_flags = torch.tensor([flags.data_ptr()], dtype=torch.int64, device="cuda:0")
    want = torch.zeros_like(src)
    want[:, 5:9, 0:7, 4:10] = src[:, 3:7, 2:9, 1:7]
    for seq in (1, 2, 3):
        dst.zero_()
        capi.check(L.pb2_peer_handshake(peer_flags.data_ptr(), flags.data_ptr(), peers.data_ptr(),
                                        1, 0, 1, seq, None))
        capi.check(L.pb2_copy_signal(t.h, counter.data_ptr(), peer_flags.data_ptr(), 1, 0, 1, seq,
                                     None))
        capi.check(L.pb2_peer_wait(flags.data_ptr(), peers.data_ptr(), 1, 1, seq, None))
        torch.cuda.synchronize()
        assert flags.tolist() == [seq, seq] and counter.item() == 0
        assert torch.equal(dst, want)
    # the copy-engine form: the caller moves the data itself, then signals
    capi.check(L.pb2_peer_handshake(peer_flags.data_ptr(), flags.data_ptr(), peers.data_ptr(),
                                    1, 0, 1, 4, None))
    capi.check(L.pb2_memcpy_d2d(dst.data_ptr(), src.data_ptr(), src.numel() * 8, None))
    capi.check(L.pb2_peer_signal(peer_flags.data_ptr(), 1, 0, 1, 4, None))
    capi.check(L.pb2_peer_wait(flags.data_ptr(), peers.data_ptr(), 1, 1, 4, None))
    torch.cuda.synchronize()
    assert flags.tolist() == [4, 4] and torch.equal(dst, src)
