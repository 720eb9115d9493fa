"""Test helpers: build C-ABI region tables for a mesh from the ORACLE's index boxes.

(The product's own table builder lives in the C++ host library and is tested separately;
these helpers let the kernels be checked in isolation through the C ABI.)
"""
import numpy as np

import oracle
from parthenon_b200 import capi


def dev_ptr(t):
    return t.data_ptr()


def strides(dims):
    nk, nj, ni = dims
    return ni, ni * nj, ni * nj * nk


def region_boxes(mesh, ir_type, prores=False):
    """[(block, nbr_index, (gid, level, ox1, ox2, ox3), s, n)] in (block, neighbor) order."""
    out = []
    for b in range(mesh.nblocks):
        for n, nb in enumerate(mesh.neighbors(b)):
            s, e = mesh.calc_indices(b, n, ir_type, prores)
            ext = tuple(e[d] - s[d] + 1 for d in range(3))
            out.append((b, n, nb, s, ext))
    return out


def match_send_region(mesh, b, nb):
    """index (in (block, neighbor) order) of the region nb's block sends to b"""
    first = np.cumsum([0] + [len(mesh.neighbors(q)) for q in range(mesh.nblocks)])
    gid, _lvl, o1, o2, o3 = nb
    for q, snb in enumerate(mesh.neighbors(gid)):
        if snb[0] == b and snb[2:] == (-o1, -o2, -o3):
            return int(first[gid] + q)
    raise AssertionError("no matching send region")


def build_bnd_tables(mesh, U_t, Uc_t, ncomp):
    """send and recv pb2_bnd_region lists whose buffer layout equals oracle.pack's"""
    loc_level = [mesh.block_loc(b)[0] for b in range(mesh.nblocks)]
    sj, sk, sc = strides(mesh.dims)
    csj, csk, csc = strides(mesh.cdims)
    bs, cbs = ncomp * sc, ncomp * csc
    send, recv = [], []
    off = 0
    offs = []
    for (b, n, nb, s, ext) in region_boxes(mesh, 0):
        coarse = nb[1] < loc_level[b]
        r = capi.BndRegion()
        r.var = (dev_ptr(Uc_t) + 8 * b * cbs) if coarse else (dev_ptr(U_t) + 8 * b * bs)
        r.buf_off = off
        r.s[:] = s
        r.n[:] = ext
        r.ncomp = ncomp
        r.stride_j, r.stride_k, r.stride_c = (csj, csk, csc) if coarse else (sj, sk, sc)
        r.flag_slot = -1
        r.status = capi.REGION_ALLOCATED
        r.value = 0.0
        send.append(r)
        offs.append(off)
        off += ncomp * ext[0] * ext[1] * ext[2]
    offs.append(off)
    for (b, n, nb, s, ext) in region_boxes(mesh, 1):
        coarse = nb[1] < loc_level[b]
        r = capi.BndRegion()
        r.var = (dev_ptr(Uc_t) + 8 * b * cbs) if coarse else (dev_ptr(U_t) + 8 * b * bs)
        r.buf_off = offs[match_send_region(mesh, b, nb)]
        r.s[:] = s
        r.n[:] = ext
        r.ncomp = ncomp
        r.stride_j, r.stride_k, r.stride_c = (csj, csk, csc) if coarse else (sj, sk, sc)
        r.flag_slot = -1
        r.status = capi.REGION_ALLOCATED | capi.REGION_BUF_ALLOCATED
        r.value = 0.0
        recv.append(r)
    return send, recv, off


def build_copy_table(mesh, U_t, ncomp):
    """fused sender-box -> receiver-box regions (uniform meshes)"""
    sj, sk, sc = strides(mesh.dims)
    bs = ncomp * sc
    sends = region_boxes(mesh, 0)
    regs = []
    for (b, n, nb, s, ext) in region_boxes(mesh, 1):
        sb, _sn, _snb, ss, sext = sends[match_send_region(mesh, b, nb)]
        assert sext == ext
        r = capi.CopyRegion()
        r.src = dev_ptr(U_t) + 8 * sb * bs
        r.dst = dev_ptr(U_t) + 8 * b * bs
        r.ss[:] = ss
        r.ds[:] = s
        r.n[:] = ext
        r.ncomp = ncomp
        r.src_stride_j, r.src_stride_k, r.src_stride_c = sj, sk, sc
        r.dst_stride_j, r.dst_stride_k, r.dst_stride_c = sj, sk, sc
        r.flag_slot = -1
        r.status = capi.REGION_ALLOCATED
        regs.append(r)
    return regs


def block_dx(mesh):
    dx = np.zeros((mesh.nblocks, 3))
    xmin = np.zeros((mesh.nblocks, 3))
    nx = [mesh._nx[0], mesh._nx[1], mesh._nx[2]]
    for b in range(mesh.nblocks):
        lo, hi = mesh.block_bounds(b)
        dx[b] = (hi - lo) / np.array(nx)
        xmin[b] = lo
    return dx, xmin


def make_geom(mesh, ncomp, dx_t):
    g = capi.PackGeom()
    g.nblocks, g.ncomp, g.ndim = mesh.nblocks, ncomp, mesh.ndim
    g.nx[:] = [int(x) for x in mesh._nx]
    g.ng = mesh.ng
    g.block_stride = ncomp * int(np.prod(mesh.dims))
    g.dx = dev_ptr(dx_t)
    return g


def refined_leaves(nrb, refine):
    """leaves of a root grid of nrb^3 blocks where the root blocks listed in `refine`
    (tuples lx1,lx2,lx3) are split once (2:1 balance is the caller's business)"""
    rl = int(np.log2(nrb))
    leaves = []
    for k in range(nrb):
        for j in range(nrb):
            for i in range(nrb):
                if (i, j, k) in refine:
                    for dk in range(2):
                        for dj in range(2):
                            for di in range(2):
                                leaves.append((rl + 1, 2 * i + di, 2 * j + dj, 2 * k + dk))
                else:
                    leaves.append((rl, i, j, k))
    return np.array(leaves, dtype=np.int32)


def leaves_from_bounds(bounds, nx_mesh, nx_block, xmin=-0.5, xmax=0.5):
    """(level, lx1, lx2, lx3) in single-tree (global) logical coordinates from the block
    bounds a reference dump records (tests/golden/refgen/burgers_dump_main.cpp).  The
    reference's own levels are tree-relative (forest.cpp:104-141), so they are rebuilt
    from the block widths instead."""
    nrb = [max(nx_mesh[d] // nx_block[d], 1) for d in range(3)]
    out = []
    for b in bounds:
        level = max(int(round(np.log2((xmax - xmin) / (b[3 + d] - b[d])))) if nrb[d] > 1 else 0
                    for d in range(3))
        lx = [int(round((b[d] - xmin) / (xmax - xmin) * (1 << level))) if nrb[d] > 1 else 0
              for d in range(3)]
        out.append((level, *lx))
    return np.array(out, dtype=np.int32), nrb


def neighbor_table(mesh):
    """[nblocks][27] int32: block owning the ghosts at receiver-side offset (ox1,ox2,ox3),
    index (ox1+1) + 3 (ox2+1) + 9 (ox3+1); -1 where there is none (uniform meshes)"""
    tab = np.full((mesh.nblocks, 27), -1, dtype=np.int32)
    for b in range(mesh.nblocks):
        for (gid, _lvl, o1, o2, o3) in mesh.neighbors(b):
            tab[b, (o1 + 1) + 3 * (o2 + 1) + 9 * (o3 + 1)] = gid
    return tab


def block_crcs(data):
    """CRC-32 of every block's bytes (the convention of tests/golden/refgen/pack_checksums.py)"""
    import zlib
    return np.array([zlib.crc32(np.ascontiguousarray(data[b]).tobytes())
                     for b in range(data.shape[0])], dtype=np.uint32)


# non-cell-centred exchange fixtures (tests/golden/refgen/tecomm_dump_main.cpp):
# (name, ndim, mesh cells per direction, block cells per direction, nghost)
TECOMM = [("tecomm_u16_b8_g2_3d", 3, 16, 8, 2), ("tecomm_u16_b8_g4_3d", 3, 16, 8, 4),
          ("tecomm_u16_b4_g2_3d", 3, 16, 4, 2), ("tecomm_u32_b8_g2_2d", 2, 32, 8, 2)]
# statically refined meshes (restriction, shared + internal prolongation): two levels in 3-D,
# three levels in 2-D
TECOMM_MULTILEVEL = [("tecomm_s16_b8_l2_3d", 3, 16, 8, 2), ("tecomm_s32_b8_l3_2d", 2, 32, 8, 2),
                     ("tecomm_s32_b8_g4_l3_2d", 2, 32, 8, 4)]
# three levels in 3-D, 197 blocks of 4^3: one CRC-32 per block and field (crc_0 / crc_1 / crc_2)
TECOMM_MULTILEVEL_CRC = [("tecomm_s16_b4_l3_3d_crc", 3, 16, 4, 2)]
# the 2-D three-level mesh with the other stock shared prolongations: (name, ..., operator id)
TECOMM_SHARED_OPS = [("tecomm_s32_b8_l3_2d_linear", 2, 32, 8, 2, 1, "linear"),
                     ("tecomm_s32_b8_l3_2d_constant", 2, 32, 8, 2, 2, "constant")]
# the same meshes with ProlongateInternalTothAndRoe registered for the face field (U_0 only)
TECOMM_TOTH_ROE = [("tecomm_s16_b8_l2_3d_tothroe", 3, 16, 8, 2),
                   ("tecomm_s32_b8_l3_2d_tothroe", 2, 32, 8, 2)]
# outflow in x1, reflecting in x2, periodic in x3 (uniform, and refined regions touching the
# boundaries)
TECOMM_BC = [("tecomm_u16_b8_g2_3d_bc", 3, 16, 8, 2), ("tecomm_s16_b8_l2_3d_bc", 3, 16, 8, 2),
             ("tecomm_s32_b8_l3_2d_bc", 2, 32, 8, 2)]
TECOMM_BC_NAMES = ("outflow", "outflow", "reflecting", "reflecting", "periodic", "periodic")
# (kind, fixture key, components): kind 1 face, 2 edge, 3 node
TECOMM_FIELDS = [(1, "U_0", 2), (2, "U_1", 1), (3, "U_2", 1)]


def tecomm_initial(nblocks, nel, ncomp, nk, nj, ni, first_gid=0):
    """the block-dependent integer code the reference-side problem generator writes into every
    entry (ghosts and shared elements included) before the exchange"""
    import numpy as np
    gid = np.arange(first_gid, first_gid + nblocks).reshape(-1, 1, 1, 1, 1, 1)
    e = np.arange(nel).reshape(1, -1, 1, 1, 1, 1)
    c = np.arange(ncomp).reshape(1, 1, -1, 1, 1, 1)
    flat = np.arange(nk * nj * ni).reshape(1, 1, 1, nk, nj, ni)
    return np.ascontiguousarray((gid + 1) * 1.0e6 + e * 1.0e5 + c * 5.0e4 + flat)


# adaptive meshes with face / edge / node fields (tests/golden/refgen/teamr_dump_main.cpp):
# (name, ndim, mesh cells, block cells, numlevel)
TEAMR = [("teamr_a32_b8_l3_2d_crc", 2, 32, 8, 3), ("teamr_a16_b4_l2_3d_crc", 3, 16, 4, 2)]
# the same runs with ProlongateInternalTothAndRoe registered for the face field
TEAMR_TOTH_ROE = [("teamr_a32_b8_l3_2d_tothroe_crc", 2, 32, 8, 3),
                  ("teamr_a16_b4_l2_3d_tothroe_crc", 3, 16, 4, 2)]
