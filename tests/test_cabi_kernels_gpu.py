"""GPU parity tests of the hand-written kernels, called through the C ABI
(include/parthenon_b200.h) and compared with the CPU oracle on the same seeded inputs.
Bit-exact for pack/unpack/copy/restrict/prolongate and for the STRICT burgers build;
<= 1e-12 relative for the FAST (FMA-contracted) burgers build."""
import ctypes as C

import numpy as np
import pytest
import torch

import oracle
from parthenon_b200 import capi
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rand_field(mesh, ncomp, seed, coarse=False):
    rng = np.random.default_rng(seed)
    shape = (mesh.nblocks, ncomp) + (mesh.cdims if coarse else mesh.dims)
    return rng.standard_normal(shape)


@pytest.mark.parametrize("ndim,nx,ng,nrb,ncomp", [
    (3, (8, 8, 8), 4, (2, 2, 2), 3),
    (3, (16, 8, 4), 2, (4, 4, 4), 2),
    (3, (7, 9, 5), 2, (2, 2, 2), 1),     # odd extents: scalar path, non-power-of-2 divisors
    (2, (16, 16), 2, (4, 4), 4),
    (3, (32, 32, 32), 4, (2, 2, 2), 11),  # burgers block shape
])
def test_pack_unpack_copy_uniform(ndim, nx, ng, nrb, ncomp):
    m = oracle.Mesh(ndim, nx, ng, nrb)
    U = rand_field(m, ncomp, 1)
    buf_ref, off = m.pack(U)
    Uref = U.copy()
    m.exchange(Uref)

    Ud = torch.from_numpy(U).to(DEV)
    dummy = torch.zeros(8, dtype=torch.float64, device=DEV)
    send, recv, total = H.build_bnd_tables(m, Ud, dummy, ncomp)
    assert total == buf_ref.size
    ts, tr = capi.Table(send, "bnd"), capi.Table(recv, "bnd")
    assert ts.elements == total and tr.elements == total
    buf = torch.full((total,), float("nan"), dtype=torch.float64, device=DEV)
    capi.check(capi.lib().pb2_pack(ts.h, buf.data_ptr(), None, None))
    torch.cuda.synchronize()
    assert np.array_equal(buf.cpu().numpy(), buf_ref)
    capi.check(capi.lib().pb2_unpack(tr.h, buf.data_ptr(), None, None))
    torch.cuda.synchronize()
    assert np.array_equal(Ud.cpu().numpy(), Uref)

    # fused same-device path
    Ud2 = torch.from_numpy(U).to(DEV)
    tc = capi.Table(H.build_copy_table(m, Ud2, ncomp), "copy")
    capi.check(capi.lib().pb2_copy(tc.h, None, None))
    torch.cuda.synchronize()
    assert np.array_equal(Ud2.cpu().numpy(), Uref)

    # descriptor-free uniform path: only a [nblocks][27] neighbour table
    Ud3 = torch.from_numpy(U).to(DEV)
    nbr = torch.from_numpy(H.neighbor_table(m)).to(DEV)
    dx = torch.ones((m.nblocks, 3), dtype=torch.float64, device=DEV)
    g = H.make_geom(m, ncomp, dx)
    capi.check(capi.lib().pb2_halo_copy_uniform(C.byref(g), Ud3.data_ptr(), nbr.data_ptr(), None))
    torch.cuda.synchronize()
    assert np.array_equal(Ud3.cpu().numpy(), Uref)


def test_unpack_through_logical_coordinate_transformation():
    """pb2_unpack of regions that carry a neighbour tree's LogicalCoordinateTransformation
    (boundary_communication.cpp:282-308): every axis permutation x flip combination, two boxes
    per table (one of them with a sign factor), against the oracle's restatement, bit for bit;
    a region without a transformation in the same table still takes the vector path"""
    import itertools
    rng = np.random.default_rng(21)
    n, ncomp = 12, 3
    s1, e1 = (2, 3, 1), (5, 4, 6)
    s2, e2 = (0, 2, 4), (4, 6, 2)
    for perm in itertools.permutations(range(3)):
        for flip in itertools.product((0, 1), repeat=3):
            U = rng.standard_normal((3, ncomp, n, n, n))
            n1, n2 = ncomp * int(np.prod(e1)), ncomp * int(np.prod(e2))
            buf = rng.standard_normal(n1 + n2 + n2)
            ref = U.copy()
            oracle.unpack_box_transformed(ref[0], s1, e1, buf[:n1], perm, flip, n, 1.0)
            oracle.unpack_box_transformed(ref[1], s2, e2, buf[n1:n1 + n2], perm, flip, n, -1.0)
            ref[2][:, s2[2]:s2[2] + e2[2], s2[1]:s2[1] + e2[1], s2[0]:s2[0] + e2[0]] = \
                buf[n1 + n2:].reshape(ncomp, e2[2], e2[1], e2[0])
            Ud = torch.from_numpy(U).to(DEV)
            regs = []
            for b, (s, e, off, fac, on) in enumerate([(s1, e1, 0, 1.0, 1), (s2, e2, n1, -1.0, 1),
                                                      (s2, e2, n1 + n2, 1.0, 0)]):
                r = capi.BndRegion()
                r.var = Ud.data_ptr() + 8 * b * ncomp * n * n * n
                r.buf_off = off
                r.s[:] = s
                r.n[:] = e
                r.ncomp = ncomp
                r.stride_j, r.stride_k, r.stride_c = n, n * n, n * n * n
                r.flag_slot = -1
                r.status = capi.REGION_ALLOCATED | capi.REGION_BUF_ALLOCATED
                r.lcoord_on = on
                r.lcoord_dir[:] = perm
                r.lcoord_flip[:] = flip
                r.lcoord_ncell = n
                r.fac = fac
                regs.append(r)
            t = capi.Table(regs, "bnd")
            bd = torch.from_numpy(buf).to(DEV)
            capi.check(capi.lib().pb2_unpack(t.h, bd.data_ptr(), None, None))
            torch.cuda.synchronize()
            assert np.array_equal(Ud.cpu().numpy(), ref), (perm, flip)


def test_halo_uniform_skips_missing_neighbours():
    """ghosts whose owner is not on this device (table entry < 0) are left untouched — they
    belong to pb2_unpack — while every other ghost is filled"""
    m = oracle.Mesh(3, (8, 8, 8), 4, (2, 2, 2))
    ncomp = 2
    U = rand_field(m, ncomp, 5)
    Uref = U.copy()
    m.exchange(Uref)
    tab = H.neighbor_table(m)
    # pretend the +x neighbour (offset index 14) of every block lives elsewhere
    tab[:, 14] = -1
    Ud = torch.from_numpy(U).to(DEV)
    nbr = torch.from_numpy(tab).to(DEV)
    dx = torch.ones((m.nblocks, 3), dtype=torch.float64, device=DEV)
    g = H.make_geom(m, ncomp, dx)
    capi.check(capi.lib().pb2_halo_copy_uniform(C.byref(g), Ud.data_ptr(), nbr.data_ptr(), None))
    torch.cuda.synchronize()
    out = Ud.cpu().numpy()
    ng, n = 4, 8
    face = (slice(None), slice(None), slice(ng, ng + n), slice(ng, ng + n), slice(ng + n, None))
    assert np.array_equal(out[face], U[face])          # untouched
    mask = np.ones(out.shape, dtype=bool)
    mask[face] = False
    assert np.array_equal(out[mask], Uref[mask])       # everything else exchanged


def test_exchange_idempotent_full_size():
    """size-independent property at benchmark block shape: a second exchange changes
    nothing, and ghost cells equal the periodic image of the interior"""
    m = oracle.Mesh(3, (32, 32, 32), 4, (4, 4, 4))
    ncomp = 11
    Ud = torch.randn((m.nblocks, ncomp) + m.dims, dtype=torch.float64, device=DEV)
    tc = capi.Table(H.build_copy_table(m, Ud, ncomp), "copy")
    capi.check(capi.lib().pb2_copy(tc.h, None, None))
    torch.cuda.synchronize()
    once = Ud.clone()
    capi.check(capi.lib().pb2_copy(tc.h, None, None))
    torch.cuda.synchronize()
    assert torch.equal(once, Ud)
    # the descriptor-free kernel on the same field: nothing changes either
    nbr = torch.from_numpy(H.neighbor_table(m)).to(DEV)
    dx = torch.ones((m.nblocks, 3), dtype=torch.float64, device=DEV)
    g = H.make_geom(m, ncomp, dx)
    capi.check(capi.lib().pb2_halo_copy_uniform(C.byref(g), Ud.data_ptr(), nbr.data_ptr(), None))
    torch.cuda.synchronize()
    assert torch.equal(once, Ud)
    # block 0 (lx=0,0,0): its -x ghost slab equals the +x interior slab of block lx=(3,0,0)
    locs = [m.block_loc(b)[1:] for b in range(m.nblocks)]
    src = locs.index((3, 0, 0))
    assert torch.equal(Ud[0, :, 4:36, 4:36, 0:4], Ud[src, :, 4:36, 4:36, 32:36])


@pytest.mark.parametrize("recon,ng", [("weno5", 4), ("linear", 2)])
@pytest.mark.parametrize("math,nscal", [("strict", 3), ("fast", 3), ("fast", 8)])
def test_burgers_stage_vs_oracle(recon, ng, math, nscal):
    """nscal = 8 is the benchmark's component count (11): the instantiations bench.py times"""
    m = oracle.Mesh(3, (16, 8, 8), ng, (2, 2, 2))
    ncomp = 3 + nscal
    B = oracle.Burgers(m, num_scalars=nscal, recon=recon)
    B.init()
    B.step()  # a developed, non-trivial state with filled ghosts
    U0 = B.U.copy()
    dt = B.dt
    # oracle: fluxes + stage 1 output (before exchange: compare interiors only)
    B.calculate_fluxes(U0)
    Fref = [B.flux(d).copy() for d in range(3)]

    dx, _ = H.block_dx(m)
    dxd = torch.from_numpy(dx).to(DEV)
    Ud = torch.from_numpy(U0).to(DEV)
    out = torch.zeros_like(Ud)
    flux = [torch.zeros_like(Ud) for _ in range(3)]
    derived = torch.zeros((m.nblocks,) + m.dims, dtype=torch.float64, device=DEV)
    dtmin = torch.full((1,), np.finfo(np.float64).max, dtype=torch.float64, device=DEV)
    a = capi.BurgersArgs()
    a.geom = H.make_geom(m, ncomp, dxd)
    a.recon = capi.RECON_WENO5 if recon == "weno5" else capi.RECON_LINEAR
    a.math = capi.MATH_STRICT if math == "strict" else capi.MATH_FAST
    a.u, a.base, a.out = Ud.data_ptr(), Ud.data_ptr(), out.data_ptr()
    for d in range(3):
        a.flux[d] = flux[d].data_ptr()
    a.derived, a.dt_min = derived.data_ptr(), dtmin.data_ptr()
    a.beta, a.dt = 1.0, dt
    # STRICT: stage = calculate_fluxes + update (stored fluxes).  FAST: stage = three
    # flux-free sweeps, so the fluxes are checked through the separate entry point.
    if math == "fast":
        capi.check(capi.lib().pb2_burgers_calculate_fluxes(C.byref(a), None))
    capi.check(capi.lib().pb2_burgers_stage(C.byref(a), None))
    torch.cuda.synchronize()

    g = m.ng
    nk, nj, ni = m.dims
    I = (slice(None), slice(None), slice(g, nk - g), slice(g, nj - g), slice(g, ni - g))
    # flux faces the reference computes: x: i in [is, ie+1]; y: j in [js, je+1]; z likewise
    Fx = (slice(None), slice(None), slice(g, nk - g), slice(g, nj - g), slice(g, ni - g + 1))
    Fy = (slice(None), slice(None), slice(g, nk - g), slice(g, nj - g + 1), slice(g, ni - g))
    Fz = (slice(None), slice(None), slice(g, nk - g + 1), slice(g, nj - g), slice(g, ni - g))
    # oracle stage-1 result
    lib = oracle.lib()
    lib.orc_burgers_stage(B.h, 1)
    # U1 is private to the oracle; recompute expected interior here from its pieces
    a1, a2, a3 = dx[:, 1] * dx[:, 2], dx[:, 0] * dx[:, 2], dx[:, 0] * dx[:, 1]
    vol = dx[:, 0] * dx[:, 1] * dx[:, 2]
    sh = (-1, 1, 1, 1, 1)
    fx, fy, fz = Fref
    du = (a1.reshape(sh) * fx[:, :, g:nk - g, g:nj - g, g + 1:ni - g + 1]
          - a1.reshape(sh) * fx[:, :, g:nk - g, g:nj - g, g:ni - g])
    du = du + (a2.reshape(sh) * fy[:, :, g:nk - g, g + 1:nj - g + 1, g:ni - g]
               - a2.reshape(sh) * fy[:, :, g:nk - g, g:nj - g, g:ni - g])
    du = du + (a3.reshape(sh) * fz[:, :, g + 1:nk - g + 1, g:nj - g, g:ni - g]
               - a3.reshape(sh) * fz[:, :, g:nk - g, g:nj - g, g:ni - g])
    dudt = -du / vol.reshape(sh)
    expect = 1.0 * (1.0 * U0[I] + 0.0 * U0[I]) + (1.0 * dt) * dudt

    got_f = [f.cpu().numpy() for f in flux]
    got = out.cpu().numpy()
    if math == "strict":
        assert np.array_equal(got_f[0][Fx], Fref[0][Fx])
        assert np.array_equal(got_f[1][Fy], Fref[1][Fy])
        assert np.array_equal(got_f[2][Fz], Fref[2][Fz])
        assert np.array_equal(got[I], expect)
    else:
        tol = 1e-12  # north_star: evolved fields within 1e-12 relative
        for gf, rf, S in zip(got_f, Fref, (Fx, Fy, Fz)):
            scale = np.abs(rf[S]).max()
            assert np.abs(gf[S] - rf[S]).max() <= tol * scale
        assert np.abs(got[I] - expect).max() <= tol * np.abs(expect).max()
    # derived and dt
    v = got[I]
    dref = 0.5 * v[:, 3] * (v[:, 0] ** 2 + v[:, 1] ** 2 + v[:, 2] ** 2)
    np.testing.assert_allclose(derived.cpu().numpy()[:, g:nk - g, g:nj - g, g:ni - g], dref,
                               rtol=1e-14, atol=0)
    inv = 1.0 / (np.abs(v[:, 0]) / dx[:, 0].reshape(-1, 1, 1, 1)
                 + np.abs(v[:, 1]) / dx[:, 1].reshape(-1, 1, 1, 1)
                 + np.abs(v[:, 2]) / dx[:, 2].reshape(-1, 1, 1, 1))
    if math == "strict":
        assert dtmin.item() == inv.min()
    else:
        assert abs(dtmin.item() - inv.min()) <= 1e-14 * inv.min()


def test_burgers_history_vs_oracle():
    m = oracle.Mesh(3, (8, 8, 8), 4, (2, 2, 2))
    B = oracle.Burgers(m, num_scalars=2)
    B.init()
    B.step()
    ref = B.history()
    dx, xmin = H.block_dx(m)
    dxd, xmd = torch.from_numpy(dx).to(DEV), torch.from_numpy(xmin).to(DEV)
    Ud = torch.from_numpy(B.U.copy()).to(DEV)
    g = H.make_geom(m, 5, dxd)
    out = np.zeros(8)
    lo, hi = np.full(3, -0.5), np.full(3, 0.5)
    dp = lambda a: a.ctypes.data_as(capi.c_double_p)
    capi.check(capi.lib().pb2_burgers_history(C.byref(g), Ud.data_ptr(), xmd.data_ptr(), dp(lo),
                                              dp(hi), dp(out), None))
    np.testing.assert_allclose(out, ref, rtol=1e-13)


def test_restrict_prolongate_multilevel():
    """3-D two-level mesh: full exchange incl. restriction on send/set and min-mod
    prolongation, bit-exact against the oracle (kernels are built with -fmad=false)."""
    nrb, nx, ng, ncomp = 2, (8, 8, 8), 2, 2
    leaves = H.refined_leaves(nrb, {(0, 0, 0)})
    m = oracle.Mesh(3, nx, ng, (nrb,) * 3, leaves=leaves)
    assert m.multilevel and m.nblocks == 15
    U = rand_field(m, ncomp, 7)
    Uc = np.zeros((m.nblocks, ncomp) + m.cdims)
    Uref, Ucref = U.copy(), Uc.copy()
    m.exchange(Uref, Ucref, prolongate=True)

    Ud, Ucd = torch.from_numpy(U).to(DEV), torch.from_numpy(Uc).to(DEV)
    send, recv, total = H.build_bnd_tables(m, Ud, Ucd, ncomp)
    lvl = [m.block_loc(b)[0] for b in range(m.nblocks)]
    sj, sk, sc = H.strides(m.dims)
    csj, csk, csc = H.strides(m.cdims)
    dx, xmin = H.block_dx(m)

    def prores(b, s, ext):
        r = capi.ProResRegion()
        r.fine = Ud.data_ptr() + 8 * b * ncomp * sc
        r.coarse = Ucd.data_ptr() + 8 * b * ncomp * csc
        r.s[:] = s
        r.n[:] = ext
        r.ncomp = ncomp
        r.fine_stride_j, r.fine_stride_k, r.fine_stride_c = sj, sk, sc
        r.coarse_stride_j, r.coarse_stride_k, r.coarse_stride_c = csj, csk, csc
        r.fine_is[:] = [ng] * 3
        r.coarse_is[:] = [ng] * 3
        r.ndim = 3
        r.status = capi.REGION_ALLOCATED
        for d in range(3):
            r.fine_dx[d] = dx[b, d]
            r.fine_xmin[d] = xmin[b, d] - ng * dx[b, d]
            r.coarse_xmin[d] = (xmin[b, d] - ng * dx[b, d]) + ng * dx[b, d] * (1 - 2)
            r.coarse_dx[d] = dx[b, d] * 2
        return r

    rs, rset, pro = [], [], []
    for (b, n, nb, s, ext) in H.region_boxes(m, 0, prores=True):
        if nb[1] < lvl[b]:
            rs.append(prores(b, s, ext))
    for b in range(m.nblocks):
        nbs = m.neighbors(b)
        restricted = any(nb[1] == lvl[b] - 1 for nb in nbs)
        for n, nb in enumerate(nbs):
            s, e = m.calc_indices(b, n, 1, True)
            ext = tuple(e[d] - s[d] + 1 for d in range(3))
            if nb[1] < lvl[b]:
                pro.append(prores(b, s, ext))
            elif restricted:
                rset.append(prores(b, s, ext))
    assert rs and rset and pro
    L = capi.lib()
    t_rs, t_rset, t_pro = (capi.Table(x, "prores") for x in (rs, rset, pro))
    ts, tr = capi.Table(send, "bnd"), capi.Table(recv, "bnd")
    buf = torch.zeros((total,), dtype=torch.float64, device=DEV)
    capi.check(L.pb2_restrict(t_rs.h, None))
    capi.check(L.pb2_pack(ts.h, buf.data_ptr(), None, None))
    capi.check(L.pb2_unpack(tr.h, buf.data_ptr(), None, None))
    capi.check(L.pb2_restrict(t_rset.h, None))
    capi.check(L.pb2_prolongate(t_pro.h, capi.PROLONG_MINMOD, None))
    torch.cuda.synchronize()
    assert np.array_equal(Ucd.cpu().numpy(), Ucref)
    assert np.array_equal(Ud.cpu().numpy(), Uref)


@pytest.mark.parametrize("ndim,nx,nrb,refine", [
    (3, (8, 8, 8), 2, {(0, 0, 0)}),
    (3, (8, 4, 6), 2, {(1, 0, 1)}),
    (2, (8, 8, 1), 4, {(1, 1, 0), (2, 1, 0)}),
])
def test_flux_correct_multilevel(ndim, nx, nrb, refine):
    """pb2_flux_correct (restrict the finer block's face fluxes straight into the coarser
    block's flux array, or into a slab + pb2_unpack) against the oracle, bit-exact."""
    ng, ncomp = 2, 3
    if ndim == 3:
        leaves = H.refined_leaves(nrb, refine)
    else:
        rl = int(np.log2(nrb))
        leaves = []
        for j in range(nrb):
            for i in range(nrb):
                if (i, j, 0) in refine:
                    leaves += [(rl + 1, 2 * i + di, 2 * j + dj, 0) for dj in range(2) for di in range(2)]
                else:
                    leaves.append((rl, i, j, 0))
        leaves = np.array(leaves, dtype=np.int32)
    m = oracle.Mesh(ndim, nx[:ndim], ng, (nrb,) * ndim, leaves=leaves)
    assert m.multilevel
    F = [rand_field(m, ncomp, 11 + d) for d in range(3)]
    Fref = [f.copy() for f in F]
    moved = m.flux_correct(Fref)
    assert moved > 0 and any(not np.array_equal(a, b) for a, b in zip(F, Fref))

    lvl = [m.block_loc(b)[0] for b in range(m.nblocks)]
    sj, sk, sc = H.strides(m.dims)
    dx, _ = H.block_dx(m)
    is_ = [ng if d < ndim else 0 for d in range(3)]

    def regions(Fd, slab_mode):
        regs, unpacks, off = [], [], 0
        for b in range(m.nblocks):
            for n, nb in enumerate(m.neighbors(b)):
                gid, nlvl, o = nb[0], nb[1], nb[2:]
                if sum(1 for x in o if x) != 1 or nlvl != lvl[b] + 1:
                    continue
                dir_ = [i for i, x in enumerate(o) if x][0]
                rs, re = m.calc_indices_flux(b, n)  # receiver box (fine index space of b)
                sn = [q for q, snb in enumerate(m.neighbors(gid))
                      if snb[0] == b and snb[2:] == tuple(-x for x in o)][0]
                ss, se = m.calc_indices_flux(gid, sn)  # sender box (coarse index space)
                ext = [re[d] - rs[d] + 1 for d in range(3)]
                assert ext == [se[d] - ss[d] + 1 for d in range(3)] and ext[dir_] == 1
                r = capi.FlxCorRegion()
                r.fine = Fd[dir_].data_ptr() + 8 * gid * ncomp * sc
                r.coarse = None if slab_mode else Fd[dir_].data_ptr() + 8 * b * ncomp * sc
                r.buf_off = off
                r.dir, r.ndim = dir_, ndim
                r.fs[:] = [(ss[d] - is_[d]) * 2 + is_[d] if d < ndim else 0 for d in range(3)]
                r.ds[:] = rs
                r.n[:] = ext
                r.ncomp = ncomp
                r.fine_stride_j, r.fine_stride_k, r.fine_stride_c = sj, sk, sc
                r.coarse_stride_j, r.coarse_stride_k, r.coarse_stride_c = sj, sk, sc
                r.status = capi.REGION_ALLOCATED
                a = [dx[gid, 1] * dx[gid, 2], dx[gid, 0] * dx[gid, 2], dx[gid, 0] * dx[gid, 1]]
                r.area = a[dir_]
                regs.append(r)
                u = capi.BndRegion()
                u.var = Fd[dir_].data_ptr() + 8 * b * ncomp * sc
                u.buf_off = off
                u.s[:] = rs
                u.n[:] = ext
                u.ncomp = ncomp
                u.stride_j, u.stride_k, u.stride_c = sj, sk, sc
                u.flag_slot = -1
                u.status = capi.REGION_ALLOCATED | capi.REGION_BUF_ALLOCATED
                unpacks.append(u)
                off += ncomp * ext[0] * ext[1] * ext[2]
        return regs, unpacks, off

    L = capi.lib()
    # fused same-device delivery
    Fd = [torch.from_numpy(f).to(DEV) for f in F]
    regs, _, total = regions(Fd, False)
    assert total == moved
    t = capi.Table(regs, "flxcor")
    assert t.elements == moved
    capi.check(L.pb2_flux_correct(t.h, None, None))
    torch.cuda.synchronize()
    for d in range(3):
        assert np.array_equal(Fd[d].cpu().numpy(), Fref[d]), d
    # slab path: restrict into a buffer, unpack on the "other device"
    Fd = [torch.from_numpy(f).to(DEV) for f in F]
    regs, unpacks, total = regions(Fd, True)
    slab = torch.full((total,), float("nan"), dtype=torch.float64, device=DEV)
    t, tu = capi.Table(regs, "flxcor"), capi.Table(unpacks, "bnd")
    capi.check(L.pb2_flux_correct(t.h, slab.data_ptr(), None))
    capi.check(L.pb2_unpack(tu.h, slab.data_ptr(), None, None))
    torch.cuda.synchronize()
    for d in range(3):
        assert np.array_equal(Fd[d].cpu().numpy(), Fref[d]), d


def test_apply_bcs_outflow_reflect():
    """pb2_apply_bcs against the oracle (outflow / reflect, faces applied in order so edges and
    corners outside the mesh are right), plus the sign flip of a vector's normal component"""
    bcs = ("outflow", "reflecting", "reflecting", "outflow", "reflecting", "reflecting")
    m = oracle.Mesh(3, (8, 6, 4), 3, (2, 2, 2), bcs=bcs)
    ncomp, ng = 3, 3
    U = rand_field(m, ncomp, 21)
    Uref = U.copy()
    m.apply_bcs(Uref)
    assert not np.array_equal(U, Uref)
    Ud = torch.from_numpy(U).to(DEV)
    nk, nj, ni = m.dims
    sj, sk, sc = H.strides(m.dims)
    code = {"outflow": 0, "reflecting": 1}

    def tables(flip):
        per_dir = [[], [], []]
        for b in range(m.nblocks):
            loc = m.block_loc(b)
            for face in range(6):
                d, inner = face // 2, face % 2 == 0
                if loc[1 + d] != (0 if inner else 1):
                    continue
                r = capi.BcRegion()
                r.var = Ud.data_ptr() + 8 * b * ncomp * sc
                r.face, r.type, r.ncomp = face, code[bcs[face]], ncomp
                r.n[:] = [ni, nj, nk]
                r.is_, r.ie = ng, ng + (8, 6, 4)[d] - 1
                r.stride_c = sc
                r.flip_mask = (1 << d) if flip else 0
                per_dir[d].append(r)
        return [capi.Table(x, "bc") for x in per_dir]

    L = capi.lib()
    for t in tables(False):
        capi.check(L.pb2_apply_bcs(t.h, None))
    torch.cuda.synchronize()
    assert np.array_equal(Ud.cpu().numpy(), Uref)
    # vector field (components = x1, x2, x3): reflecting faces flip the normal component
    V = rand_field(m, ncomp, 22)
    Vd = torch.from_numpy(V).to(DEV)
    Ud = Vd
    for t in tables(True):
        capi.check(L.pb2_apply_bcs(t.h, None))
    torch.cuda.synchronize()
    out = Vd.cpu().numpy()
    b = 0  # block (0,0,0): inner faces; ix1 outflow (no flip), ix2 / ix3 reflecting
    assert np.array_equal(out[b, :, ng:-ng, ng:-ng, :ng],
                          np.repeat(V[b, :, ng:-ng, ng:-ng, ng:ng + 1], ng, axis=-1))
    mirror = V[b, :, ng:-ng, ng:2 * ng, ng:-ng][:, :, ::-1, :]
    sign = np.array([1.0, -1.0, 1.0])[:, None, None, None]
    assert np.array_equal(out[b, :, ng:-ng, :ng, ng:-ng], sign * mirror)


def test_weighted_sum_and_flux_div():
    n = 100003
    x = torch.randn(n, dtype=torch.float64, device=DEV)
    y = torch.randn(n, dtype=torch.float64, device=DEV)
    z = torch.empty_like(x)
    capi.check(capi.lib().pb2_weighted_sum(x.data_ptr(), y.data_ptr(), 0.5, 0.25, z.data_ptr(),
                                           n, None))
    torch.cuda.synchronize()
    assert torch.equal(z, 0.5 * x + 0.25 * y)


@pytest.mark.parametrize("shape,nrb,recon,ng", [((32, 32, 32), (2, 2, 2), "weno5", 4),
                                                ((16, 8, 8), (2, 2, 2), "weno5", 4),
                                                ((8, 8, 8), (1, 1, 1), "weno5", 4),
                                                ((16, 12, 1), (4, 4, 1), "linear", 2)])
def test_fast_stage_reads_neighbours_instead_of_ghosts(shape, nrb, recon, ng):
    """pb2_burgers_args::nbr_direct: stencil values beyond a block face come from the interior
    of the same-device neighbour.  The stage on a field whose ghost cells are NaN must equal the
    stage on the exchanged field bit for bit — at the benchmark's block shape (32^3, 11
    components: the compile-time-geometry kernels), on generic shapes, on a single block that is
    its own periodic neighbour 26 times, and in 2-D."""
    ndim = 3 if shape[2] > 1 else 2
    m = oracle.Mesh(ndim, shape, ng, nrb)
    nscal = 8
    ncomp = 3 + nscal
    B = oracle.Burgers(m, num_scalars=nscal, recon=recon)
    B.init()
    U0 = B.U.copy()  # ghosts exchanged
    g = m.ng
    nk, nj, ni = m.dims
    Unan = np.full_like(U0, np.nan)
    I = (slice(None), slice(None), slice(g, nk - g) if ndim > 2 else slice(None),
         slice(g, nj - g), slice(g, ni - g))
    Unan[I] = U0[I]
    dx, _ = H.block_dx(m)
    dxd = torch.from_numpy(dx).to(DEV)
    nbr = torch.from_numpy(H.neighbor_table(m)).to(DEV)
    outs = []
    for direct in (False, True):
        Ud = torch.from_numpy(Unan if direct else U0).to(DEV)
        out = torch.zeros_like(Ud)
        derived = torch.zeros((m.nblocks,) + m.dims, dtype=torch.float64, device=DEV)
        dtmin = torch.full((1,), np.finfo(np.float64).max, dtype=torch.float64, device=DEV)
        a = capi.BurgersArgs()
        a.geom = H.make_geom(m, ncomp, dxd)
        a.recon = capi.RECON_WENO5 if recon == "weno5" else capi.RECON_LINEAR
        a.math = capi.MATH_FAST
        a.u, a.base, a.out = Ud.data_ptr(), Ud.data_ptr(), out.data_ptr()
        a.derived, a.dt_min = derived.data_ptr(), dtmin.data_ptr()
        a.beta, a.dt = 1.0, B.dt
        if direct:
            a.nbr_direct = nbr.data_ptr()
        capi.check(capi.lib().pb2_burgers_stage(C.byref(a), None))
        torch.cuda.synchronize()
        outs.append((out.cpu().numpy()[I], derived.cpu().numpy(), float(dtmin.item())))
    assert np.all(np.isfinite(outs[1][0]))
    assert np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][1], outs[1][1]) and outs[0][2] == outs[1][2]
