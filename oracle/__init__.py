"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE ONLY — see pb2_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libpb2_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "pb2_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        L = _lib
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        L.orc_mesh_create.restype = C.c_void_p
        L.orc_mesh_create.argtypes = [C.c_int, ip, C.c_int, ip, dp, dp, C.c_int, ip]
        L.orc_mesh_create_uniform.restype = C.c_void_p
        L.orc_mesh_create_uniform.argtypes = [C.c_int, ip, C.c_int, ip, dp, dp]
        L.orc_mesh_destroy.argtypes = [C.c_void_p]
        L.orc_mesh_set_next_bcs.argtypes = [ip]
        L.orc_apply_bcs.argtypes = [C.c_void_p, dp, C.c_int]
        L.orc_mesh_nblocks.argtypes = [C.c_void_p]
        L.orc_mesh_multilevel.argtypes = [C.c_void_p]
        L.orc_mesh_dims.argtypes = [C.c_void_p, ip, ip]
        L.orc_mesh_block_loc.argtypes = [C.c_void_p, C.c_int, ip]
        L.orc_mesh_block_bounds.argtypes = [C.c_void_p, C.c_int, dp, dp]
        L.orc_mesh_num_neighbors.argtypes = [C.c_void_p, C.c_int]
        L.orc_mesh_neighbor.argtypes = [C.c_void_p, C.c_int, C.c_int, ip]
        L.orc_calc_indices.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, ip, ip]
        L.orc_exchange.restype = C.c_int64
        L.orc_exchange.argtypes = [C.c_void_p, dp, dp, C.c_int, C.c_int]
        L.orc_count_regions.restype = C.c_int64
        L.orc_count_regions.argtypes = [C.c_void_p]
        L.orc_pack.restype = C.c_int64
        L.orc_pack.argtypes = [C.c_void_p, dp, dp, C.c_int, dp, C.POINTER(C.c_int64)]
        L.orc_unpack.argtypes = [C.c_void_p, dp, dp, C.c_int, dp, C.POINTER(C.c_int64)]
        L.orc_restrict_send.argtypes = [C.c_void_p, dp, dp, C.c_int]
        L.orc_restrict_set.argtypes = [C.c_void_p, dp, dp, C.c_int]
        L.orc_prolongate.argtypes = [C.c_void_p, dp, dp, C.c_int]
        L.orc_calc_indices_flux.argtypes = [C.c_void_p, C.c_int, C.c_int, ip, ip]
        L.orc_te_extents.argtypes = [C.c_void_p, C.c_int, ip]
        L.orc_calc_indices_te.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          ip, ip]
        L.orc_te_recv_mask.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, ip]
        L.orc_exchange_te.argtypes = [C.c_void_p, dp, C.c_int, C.c_int]
        L.orc_exchange_te.restype = C.c_int64
        L.orc_exchange_te_ml.argtypes = [C.c_void_p, dp, dp, C.c_int, C.c_int]
        L.orc_exchange_te_ml.restype = C.c_int64
        L.orc_exchange_te_ml_op.argtypes = [C.c_void_p, dp, dp, C.c_int, C.c_int, C.c_int]
        L.orc_exchange_te_ml_op.restype = C.c_int64
        L.orc_calc_indices_te_general.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                                  C.c_int, C.c_int, ip, ip, ip]
        L.orc_exchange_te_ml_toth_roe.argtypes = [C.c_void_p, dp, dp, C.c_int]
        L.orc_exchange_te_ml_toth_roe.restype = C.c_int64
        L.orc_flux_correct_edge.argtypes = [C.c_void_p, dp, dp, C.c_int, C.c_void_p]
        L.orc_flux_correct_edge.restype = C.c_int64
        L.orc_flux_correct.restype = C.c_int64
        L.orc_flux_correct.argtypes = [C.c_void_p, C.POINTER(dp), C.c_int]
        L.orc_weno5z.argtypes = [C.c_double] * 5 + [dp, dp]
        L.orc_linear.argtypes = [C.c_double] * 3 + [dp, dp]
        L.orc_lr_to_flux.argtypes = [C.c_double] * 8 + [dp] * 5
        L.orc_burgers_ic.argtypes = [C.c_void_p, dp, C.c_int]
        L.orc_burgers_create.restype = C.c_void_p
        L.orc_burgers_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
        L.orc_burgers_destroy.argtypes = [C.c_void_p]
        for f in ("orc_burgers_U", "orc_burgers_derived"):
            getattr(L, f).restype = dp
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_burgers_flux.restype = dp
        L.orc_burgers_flux.argtypes = [C.c_void_p, C.c_int]
        L.orc_burgers_init.argtypes = [C.c_void_p]
        L.orc_burgers_dt.restype = C.c_double
        L.orc_burgers_dt.argtypes = [C.c_void_p]
        L.orc_burgers_time.restype = C.c_double
        L.orc_burgers_time.argtypes = [C.c_void_p]
        L.orc_burgers_step.argtypes = [C.c_void_p]
        L.orc_burgers_calculate_fluxes.argtypes = [C.c_void_p, dp]
        L.orc_burgers_stage.argtypes = [C.c_void_p, C.c_int]
        L.orc_burgers_history.argtypes = [C.c_void_p, dp]
        L.orc_burgers_cycle.argtypes = [C.c_void_p]
        L.orc_advection_create.restype = C.c_void_p
        L.orc_advection_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, dp,
                                           C.c_double]
        L.orc_advection_destroy.argtypes = [C.c_void_p]
        L.orc_advection_U.restype = dp
        L.orc_advection_U.argtypes = [C.c_void_p]
        L.orc_advection_flux.restype = dp
        L.orc_advection_flux.argtypes = [C.c_void_p, C.c_int]
        L.orc_advection_init.argtypes = [C.c_void_p]
        L.orc_advection_step.argtypes = [C.c_void_p]
        L.orc_advection_stage.argtypes = [C.c_void_p, C.c_int]
        L.orc_advection_calculate_fluxes.argtypes = [C.c_void_p, dp]
        L.orc_advection_dt.restype = C.c_double
        L.orc_advection_dt.argtypes = [C.c_void_p]
        L.orc_advection_time.restype = C.c_double
        L.orc_advection_time.argtypes = [C.c_void_p]
        L.orc_sparse_create.restype = C.c_void_p
        L.orc_sparse_create.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double,
                                        C.c_double, C.c_int]
        L.orc_sparse_destroy.argtypes = [C.c_void_p]
        L.orc_sparse_U.restype = dp
        L.orc_sparse_U.argtypes = [C.c_void_p, C.c_int]
        L.orc_sparse_alloc.restype = C.POINTER(C.c_ubyte)
        L.orc_sparse_alloc.argtypes = [C.c_void_p]
        L.orc_sparse_init.argtypes = [C.c_void_p]
        L.orc_sparse_step.argtypes = [C.c_void_p]
        L.orc_sparse_dt.restype = C.c_double
        L.orc_sparse_dt.argtypes = [C.c_void_p]
        L.orc_sparse_time.restype = C.c_double
        L.orc_sparse_time.argtypes = [C.c_void_p]
        L.orc_amr_create.restype = C.c_void_p
        L.orc_amr_create.argtypes = [C.c_int, ip, C.c_int, ip, dp, dp, C.c_int, C.c_int,
                                     C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, dp,
                                     C.c_double]
        L.orc_amr_create_burgers.restype = C.c_void_p
        L.orc_amr_create_burgers.argtypes = [C.c_int, ip, C.c_int, ip, dp, dp, C.c_int, C.c_int,
                                             C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                             C.c_double]
        L.orc_amr_create_te.argtypes = [C.c_int, ip, C.c_int, ip, dp, dp, C.c_int, C.c_int]
        L.orc_amr_create_te.restype = C.c_void_p
        L.orc_amr_te_cycle.argtypes = [C.c_void_p, C.c_int]
        L.orc_amr_te_field.argtypes = [C.c_void_p, C.c_int]
        L.orc_amr_te_field.restype = dp
        L.orc_amr_te_set_toth_roe.argtypes = [C.c_void_p, C.c_int]
        L.orc_amr_create_sparse.argtypes = [C.c_int, ip, C.c_int, ip, dp, dp, C.c_int, C.c_int,
                                            C.c_double, C.c_double, C.c_double, C.c_double,
                                            C.c_double, C.c_double, C.c_int]
        L.orc_amr_create_sparse.restype = C.c_void_p
        for fn in ("orc_amr_sparse_init", "orc_amr_sparse_step"):
            getattr(L, fn).argtypes = [C.c_void_p]
        L.orc_amr_sparse_regrid.argtypes = [C.c_void_p]
        L.orc_amr_sparse_state.argtypes = [C.c_void_p]
        L.orc_amr_sparse_state.restype = C.c_void_p
        L.orc_amr_tags.argtypes = [C.c_void_p, ip]
        L.orc_amr_deref_counts.argtypes = [C.c_void_p, ip]
        L.orc_amr_destroy.argtypes = [C.c_void_p]
        L.orc_amr_mesh.restype = C.c_void_p
        L.orc_amr_mesh.argtypes = [C.c_void_p]
        L.orc_amr_U.restype = dp
        L.orc_amr_U.argtypes = [C.c_void_p]
        L.orc_amr_init.argtypes = [C.c_void_p]
        L.orc_amr_step.argtypes = [C.c_void_p]
        L.orc_amr_regrid.argtypes = [C.c_void_p]
        L.orc_amr_dt.restype = C.c_double
        L.orc_amr_dt.argtypes = [C.c_void_p]
        L.orc_amr_time.restype = C.c_double
        L.orc_amr_time.argtypes = [C.c_void_p]
        L.orc_set_num_threads.argtypes = [C.c_int]
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class Mesh:
    """Single-tree block-structured mesh; leaves=None => uniform root grid."""

    BCS = {"periodic": 0, "outflow": 1, "reflecting": 2}

    def __init__(self, ndim, nx, ng, nrb, xmin=(-0.5,) * 3, xmax=(0.5,) * 3, leaves=None,
                 bcs=None):
        """bcs: six names (ix1, ox1, ix2, ox2, ix3, ox3) from BCS, default periodic"""
        L = lib()
        self.bcs = tuple(bcs) if bcs else ("periodic",) * 6
        bc = np.array([self.BCS[b] for b in self.bcs], dtype=np.int32)
        L.orc_mesh_set_next_bcs(_ip(bc))
        self.ndim, self.ng = ndim, ng
        self._nx = np.array(list(nx) + [1] * (3 - len(nx)), dtype=np.int32)
        self._nrb = np.array(list(nrb) + [1] * (3 - len(nrb)), dtype=np.int32)
        self._xmin = np.array(xmin, dtype=np.float64)
        self._xmax = np.array(xmax, dtype=np.float64)
        if leaves is None:
            self.h = L.orc_mesh_create_uniform(ndim, _ip(self._nx), ng, _ip(self._nrb),
                                               _dp(self._xmin), _dp(self._xmax))
        else:
            lv = np.ascontiguousarray(leaves, dtype=np.int32)
            self.h = L.orc_mesh_create(ndim, _ip(self._nx), ng, _ip(self._nrb),
                                       _dp(self._xmin), _dp(self._xmax), lv.shape[0], _ip(lv))
        L.orc_mesh_set_next_bcs(None)
        if not self.h:
            raise ValueError("unsupported mesh")
        self.nblocks = L.orc_mesh_nblocks(self.h)
        self.multilevel = bool(L.orc_mesh_multilevel(self.h))
        d = np.zeros(3, dtype=np.int32)
        c = np.zeros(3, dtype=np.int32)
        L.orc_mesh_dims(self.h, _ip(d), _ip(c))
        self.dims = tuple(int(x) for x in d)    # (nk, nj, ni)
        self.cdims = tuple(int(x) for x in c)

    @classmethod
    def from_handle(cls, h, ndim=2):
        """wrap a mesh the C oracle already built (oracle/forest.py: orc_mesh_create_custom)"""
        self = cls.__new__(cls)
        L = lib()
        self.h = h
        self.bcs = None
        self.ndim = ndim
        self.nblocks = L.orc_mesh_nblocks(self.h)
        self.multilevel = bool(L.orc_mesh_multilevel(self.h))
        d = np.zeros(3, dtype=np.int32)
        c = np.zeros(3, dtype=np.int32)
        L.orc_mesh_dims(self.h, _ip(d), _ip(c))
        self.dims = tuple(int(x) for x in d)
        self.cdims = tuple(int(x) for x in c)
        return self

    def __del__(self):
        try:
            lib().orc_mesh_destroy(self.h)
        except Exception:
            pass

    def block_loc(self, b):
        o = np.zeros(4, dtype=np.int32)
        lib().orc_mesh_block_loc(self.h, b, _ip(o))
        return tuple(int(x) for x in o)

    def block_bounds(self, b):
        lo = np.zeros(3)
        hi = np.zeros(3)
        lib().orc_mesh_block_bounds(self.h, b, _dp(lo), _dp(hi))
        return lo, hi

    def neighbors(self, b):
        out = []
        o = np.zeros(5, dtype=np.int32)
        for n in range(lib().orc_mesh_num_neighbors(self.h, b)):
            lib().orc_mesh_neighbor(self.h, b, n, _ip(o))
            out.append(tuple(int(x) for x in o))
        return out

    def calc_indices(self, b, n, ir_type, prores=False):
        s = np.zeros(3, dtype=np.int32)
        e = np.zeros(3, dtype=np.int32)
        lib().orc_calc_indices(self.h, b, n, ir_type, int(prores), _ip(s), _ip(e))
        return tuple(int(x) for x in s), tuple(int(x) for x in e)

    def field(self, ncomp):
        return np.zeros((self.nblocks, ncomp) + self.dims)

    def coarse_field(self, ncomp):
        return np.zeros((self.nblocks, ncomp) + self.cdims)

    def exchange(self, U, Uc=None, prolongate=False):
        assert U.flags.c_contiguous
        return lib().orc_exchange(self.h, _dp(U), _dp(Uc) if Uc is not None else None,
                                  U.shape[1], int(prolongate))

    def pack(self, U, Uc=None):
        nreg = lib().orc_count_regions(self.h)
        off = np.zeros(nreg + 1, dtype=np.int64)
        offp = off.ctypes.data_as(C.POINTER(C.c_int64))
        ucp = _dp(Uc) if Uc is not None else None
        total = lib().orc_pack(self.h, _dp(U), ucp, U.shape[1], None, offp)
        buf = np.zeros(total)
        lib().orc_pack(self.h, _dp(U), ucp, U.shape[1], _dp(buf), offp)
        return buf, off

    def unpack(self, U, buf, off, Uc=None):
        lib().orc_unpack(self.h, _dp(U), _dp(Uc) if Uc is not None else None, U.shape[1],
                         _dp(buf), off.ctypes.data_as(C.POINTER(C.c_int64)))

    def calc_indices_flux(self, b, n):
        s = np.zeros(3, dtype=np.int32)
        e = np.zeros(3, dtype=np.int32)
        lib().orc_calc_indices_flux(self.h, b, n, _ip(s), _ip(e))
        return tuple(int(x) for x in s), tuple(int(x) for x in e)

    def apply_bcs(self, U):
        lib().orc_apply_bcs(self.h, _dp(U), U.shape[1])

    # non-cell-centred fields (uniform meshes): kind 0 cell, 1 face, 2 edge, 3 node
    def te_extents(self, kind):
        pn = np.zeros(3, dtype=np.int32)
        lib().orc_te_extents(self.h, kind, _ip(pn))
        return tuple(int(x) for x in pn[::-1])  # (nk, nj, ni)

    def calc_indices_te(self, b, n, kind, el, ir_type):
        s = np.zeros(3, dtype=np.int32)
        e = np.zeros(3, dtype=np.int32)
        lib().orc_calc_indices_te(self.h, b, n, kind, el, ir_type, _ip(s), _ip(e))
        return tuple(int(x) for x in s), tuple(int(x) for x in e)

    def calc_indices_te_general(self, b, n, kind, el, ir_type, prores=False):
        """(s, e, mask[k][j][i]) of region (b, n), element (kind, el), on any mesh"""
        s = np.zeros(3, dtype=np.int32)
        e = np.zeros(3, dtype=np.int32)
        mk = np.zeros(27, dtype=np.int32)
        lib().orc_calc_indices_te_general(self.h, b, n, kind, el, ir_type, int(prores), _ip(s),
                                          _ip(e), _ip(mk))
        return tuple(int(x) for x in s), tuple(int(x) for x in e), mk.reshape(3, 3, 3)

    def te_recv_mask(self, b, n, kind, el):
        """[k][j][i] over (-1, 0, 1): which entries of the receive box are written"""
        mk = np.zeros(27, dtype=np.int32)
        lib().orc_te_recv_mask(self.h, b, n, kind, el, _ip(mk))
        return mk.reshape(3, 3, 3)

    def exchange_te(self, U, kind, toth_roe=False, shared_op=0):
        """U: [nblocks][elements][ncomp][nk'][nj'][ni'] (te_extents), exchanged in place;
        multilevel meshes get scratch coarse buffers (restriction / prolongation included)"""
        assert U.flags.c_contiguous and U.shape[3:] == self.te_extents(kind)
        if not self.multilevel:
            return lib().orc_exchange_te(self.h, _dp(U), U.shape[2], kind)
        cd = tuple(n + (1 if (kind != 0 and n > 1) else 0) for n in self.cdims)
        Uc = np.zeros(U.shape[:3] + cd)
        if toth_roe:
            assert kind == 1, "Toth & Roe prolongation is defined for face fields"
            return lib().orc_exchange_te_ml_toth_roe(self.h, _dp(U), _dp(Uc), U.shape[2])
        if shared_op:
            return lib().orc_exchange_te_ml_op(self.h, _dp(U), _dp(Uc), U.shape[2], kind, shared_op)
        return lib().orc_exchange_te_ml(self.h, _dp(U), _dp(Uc), U.shape[2], kind)

    def flux_correct_edge(self, F, deliveries=None):
        """flux correction of a FACE field: F is its edge-centred flux field
        [nblocks][3 elements][ncomp][nk'][nj'][ni'] (te_extents(2)), corrected in place; returns
        the number of Reals the fine blocks sent.  deliveries: optional int32 array of F's shape
        that receives how often each entry was written (> 1: the reference is not deterministic
        there, see pb2_oracle.c)"""
        assert F.flags.c_contiguous and F.shape[1] == 3 and F.shape[3:] == self.te_extents(2)
        cd = tuple(n + (1 if n > 1 else 0) for n in self.cdims)
        Fc = np.zeros(F.shape[:3] + cd)
        w = None
        if deliveries is not None:
            assert deliveries.dtype == np.int32 and deliveries.shape == F.shape
            assert deliveries.flags.c_contiguous
            w = deliveries.ctypes.data
        return lib().orc_flux_correct_edge(self.h, _dp(F), _dp(Fc), F.shape[2], w)

    def flux_correct(self, F):
        """F: three face-flux arrays [nblocks][ncomp][nk][nj][ni], corrected in place"""
        dp = C.POINTER(C.c_double)
        arr = (dp * 3)(*[_dp(f) for f in F])
        return lib().orc_flux_correct(self.h, arr, F[0].shape[1])

    def burgers_ic(self, U):
        lib().orc_burgers_ic(self.h, _dp(U), U.shape[1])


class Burgers:
    def __init__(self, mesh, num_scalars=8, recon="weno5", cfl=0.8):
        self.mesh = mesh
        self.ncomp = 3 + num_scalars
        self.h = lib().orc_burgers_create(mesh.h, num_scalars, 0 if recon == "weno5" else 1, cfl)

    def __del__(self):
        try:
            lib().orc_burgers_destroy(self.h)
        except Exception:
            pass

    def _view(self, p, shape):
        n = int(np.prod(shape))
        return np.ctypeslib.as_array(p, shape=(n,)).reshape(shape)

    @property
    def U(self):
        return self._view(lib().orc_burgers_U(self.h),
                          (self.mesh.nblocks, self.ncomp) + self.mesh.dims)

    @property
    def derived(self):
        return self._view(lib().orc_burgers_derived(self.h), (self.mesh.nblocks,) + self.mesh.dims)

    def flux(self, d):
        return self._view(lib().orc_burgers_flux(self.h, d),
                          (self.mesh.nblocks, self.ncomp) + self.mesh.dims)

    def init(self):
        lib().orc_burgers_init(self.h)

    def step(self):
        lib().orc_burgers_step(self.h)

    def calculate_fluxes(self, U):
        lib().orc_burgers_calculate_fluxes(self.h, _dp(U))

    @property
    def dt(self):
        return lib().orc_burgers_dt(self.h)

    @property
    def time(self):
        return lib().orc_burgers_time(self.h)

    def history(self):
        o = np.zeros(8)
        lib().orc_burgers_history(self.h, _dp(o))
        return o


class Advection:
    """example/advection with a constant velocity; profile 'smooth_gaussian' | 'hard_sphere'"""

    PROFILES = {"smooth_gaussian": 1, "hard_sphere": 2}

    def __init__(self, mesh, vec_size=1, profile="hard_sphere", amp=1e-6, v=(1.0, 1.0, 1.0),
                 cfl=0.45):
        self.mesh = mesh
        self.ncomp = vec_size
        self._v = np.array(v, dtype=np.float64)
        self.h = lib().orc_advection_create(mesh.h, vec_size, self.PROFILES[profile], amp,
                                            _dp(self._v), cfl)

    def __del__(self):
        try:
            lib().orc_advection_destroy(self.h)
        except Exception:
            pass

    def _view(self, p):
        shape = (self.mesh.nblocks, self.ncomp) + self.mesh.dims
        return np.ctypeslib.as_array(p, shape=(int(np.prod(shape)),)).reshape(shape)

    @property
    def U(self):
        return self._view(lib().orc_advection_U(self.h))

    def flux(self, d):
        return self._view(lib().orc_advection_flux(self.h, d))

    def init(self):
        lib().orc_advection_init(self.h)

    def step(self):
        lib().orc_advection_step(self.h)

    def calculate_fluxes(self, U):
        lib().orc_advection_calculate_fluxes(self.h, _dp(U))

    @property
    def dt(self):
        return lib().orc_advection_dt(self.h)

    @property
    def time(self):
        return lib().orc_advection_time(self.h)


class AmrAdvection:
    """example/advection with refinement = adaptive: the mesh is owned by the C side and
    changes as blocks are refined / derefined"""

    def __init__(self, ndim, nx, ng, nrb, numlevel, derefine_count=10, refine_tol=0.3,
                 derefine_tol=0.03, vec_size=1, profile="hard_sphere", amp=1e-6,
                 v=(1.0, 1.0, 1.0), cfl=0.45, xmin=(-0.5,) * 3, xmax=(0.5,) * 3, burgers=None):
        """burgers: None (example/advection) or dict(num_scalars=, recon=, vector_i=) for
        benchmarks/burgers with the deck's derivative_order_1 criterion"""
        L = lib()
        self.ndim, self.ncomp = ndim, vec_size
        if burgers is not None:
            nx3 = np.array(list(nx) + [1] * (3 - len(nx)), dtype=np.int32)
            nrb3 = np.array(list(nrb) + [1] * (3 - len(nrb)), dtype=np.int32)
            lo, hi = np.array(xmin, dtype=np.float64), np.array(xmax, dtype=np.float64)
            self.ncomp = 3 + burgers["num_scalars"]
            self.h = L.orc_amr_create_burgers(
                ndim, _ip(nx3), ng, _ip(nrb3), _dp(lo), _dp(hi), numlevel, derefine_count,
                refine_tol, derefine_tol, burgers.get("vector_i", 3), burgers["num_scalars"],
                0 if burgers.get("recon", "weno5") == "weno5" else 1, cfl)
            return
        nx3 = np.array(list(nx) + [1] * (3 - len(nx)), dtype=np.int32)
        nrb3 = np.array(list(nrb) + [1] * (3 - len(nrb)), dtype=np.int32)
        vv = np.array(v, dtype=np.float64)
        lo, hi = np.array(xmin, dtype=np.float64), np.array(xmax, dtype=np.float64)
        self.h = L.orc_amr_create(ndim, _ip(nx3), ng, _ip(nrb3), _dp(lo), _dp(hi), numlevel,
                                  derefine_count, refine_tol, derefine_tol, vec_size,
                                  Advection.PROFILES[profile], amp, _dp(vv), cfl)

    def __del__(self):
        try:
            lib().orc_amr_destroy(self.h)
        except Exception:
            pass

    def init(self):
        lib().orc_amr_init(self.h)

    def step(self):
        """Step + tagging; the mesh is adapted by regrid()"""
        lib().orc_amr_step(self.h)

    def regrid(self):
        return bool(lib().orc_amr_regrid(self.h))

    @property
    def tags(self):
        out = np.zeros(self.nblocks, dtype=np.int32)
        lib().orc_amr_tags(self.h, _ip(out))
        return out

    @property
    def deref_counts(self):
        out = np.zeros(self.nblocks, dtype=np.int32)
        lib().orc_amr_deref_counts(self.h, _ip(out))
        return out

    def _mesh(self):
        return lib().orc_amr_mesh(self.h)

    @property
    def nblocks(self):
        return lib().orc_mesh_nblocks(self._mesh())

    @property
    def block_locs(self):
        L, mh = lib(), self._mesh()
        out = np.zeros((self.nblocks, 4), dtype=np.int32)
        o = np.zeros(4, dtype=np.int32)
        for b in range(out.shape[0]):
            L.orc_mesh_block_loc(mh, b, _ip(o))
            out[b] = o
        return out

    @property
    def U(self):
        L, mh = lib(), self._mesh()
        d, c = np.zeros(3, dtype=np.int32), np.zeros(3, dtype=np.int32)
        L.orc_mesh_dims(mh, _ip(d), _ip(c))
        shape = (self.nblocks, self.ncomp) + tuple(int(x) for x in d)
        return np.ctypeslib.as_array(L.orc_amr_U(self.h), shape=(int(np.prod(shape)),)).reshape(shape).copy()

    @property
    def dt(self):
        return lib().orc_amr_dt(self.h)

    @property
    def time(self):
        return lib().orc_amr_time(self.h)


class AmrNonCellCentred:
    """the adaptive face / edge / node field application of
    tests/golden/refgen/teamr_dump_main.cpp (fields never evolve, the mesh follows a moving
    geometric criterion): pins the remesh data movement of non-cell-centred fields"""

    def __init__(self, ndim, nx, ng, nrb, numlevel, derefine_count=2, xmin=(-0.5,) * 3,
                 xmax=(0.5,) * 3, toth_roe=False):
        nx3 = np.array(list(nx) + [1] * (3 - len(nx)), dtype=np.int32)
        nrb3 = np.array(list(nrb) + [1] * (3 - len(nrb)), dtype=np.int32)
        lo, hi = np.array(xmin, dtype=np.float64), np.array(xmax, dtype=np.float64)
        self.ndim = ndim
        self.h = lib().orc_amr_create_te(ndim, _ip(nx3), ng, _ip(nrb3), _dp(lo), _dp(hi), numlevel,
                                         derefine_count)
        lib().orc_amr_te_set_toth_roe(self.h, int(toth_roe))

    def __del__(self):
        try:
            lib().orc_amr_destroy(self.h)
        except Exception:
            pass

    def init(self):
        lib().orc_amr_init(self.h)

    def cycle(self, c):
        """tag with the criterion of cycle c and remesh; True if the mesh changed"""
        return bool(lib().orc_amr_te_cycle(self.h, c))

    @property
    def nblocks(self):
        return lib().orc_mesh_nblocks(lib().orc_amr_mesh(self.h))

    @property
    def block_locs(self):
        m = lib().orc_amr_mesh(self.h)
        out = np.zeros((self.nblocks, 4), dtype=np.int32)
        for b in range(self.nblocks):
            lib().orc_mesh_block_loc(m, b, _ip(out[b]))
        return out

    def field(self, f):
        """f = 0 face [nb][3*2], 1 edge [nb][3], 2 node [nb][1], each with padded extents"""
        m = lib().orc_amr_mesh(self.h)
        pn = np.zeros(3, dtype=np.int32)
        lib().orc_te_extents(m, f + 1, _ip(pn))
        ncs = (6, 3, 1)[f]
        shape = (self.nblocks, ncs, int(pn[2]), int(pn[1]), int(pn[0]))
        p = lib().orc_amr_te_field(self.h, f)
        return np.ctypeslib.as_array(p, shape=(int(np.prod(shape)),)).reshape(shape).copy()


class AmrSparseAdvection:
    """example/sparse_advection with refinement = adaptive (2-D in the reference)"""

    def __init__(self, ndim, nx, ng, nrb, numlevel, derefine_count=10, refine_tol=0.3,
                 derefine_tol=0.03, speed=1.5, cfl=0.45, alloc_threshold=1e-5,
                 dealloc_threshold=1e-6, dealloc_count=5, xmin=(-1.0,) * 3, xmax=(1.0,) * 3):
        nx3 = np.array(list(nx) + [1] * (3 - len(nx)), dtype=np.int32)
        nrb3 = np.array(list(nrb) + [1] * (3 - len(nrb)), dtype=np.int32)
        lo, hi = np.array(xmin, dtype=np.float64), np.array(xmax, dtype=np.float64)
        self.h = lib().orc_amr_create_sparse(ndim, _ip(nx3), ng, _ip(nrb3), _dp(lo), _dp(hi),
                                             numlevel, derefine_count, refine_tol, derefine_tol,
                                             speed, cfl, alloc_threshold, dealloc_threshold,
                                             dealloc_count)

    def __del__(self):
        try:
            lib().orc_amr_destroy(self.h)
        except Exception:
            pass

    def init(self):
        lib().orc_amr_sparse_init(self.h)

    def step(self):
        lib().orc_amr_sparse_step(self.h)

    def regrid(self):
        return bool(lib().orc_amr_sparse_regrid(self.h))

    @property
    def nblocks(self):
        return lib().orc_mesh_nblocks(lib().orc_amr_mesh(self.h))

    @property
    def block_locs(self):
        m = lib().orc_amr_mesh(self.h)
        out = np.zeros((self.nblocks, 4), dtype=np.int32)
        for b in range(self.nblocks):
            lib().orc_mesh_block_loc(m, b, _ip(out[b]))
        return out

    @property
    def time(self):
        return lib().orc_sparse_time(lib().orc_amr_sparse_state(self.h))

    @property
    def dt(self):
        return lib().orc_sparse_dt(lib().orc_amr_sparse_state(self.h))

    @property
    def U(self):
        """[nblocks][4][nk][nj][ni]; NaN where a field is not allocated"""
        L = lib()
        st, m = L.orc_amr_sparse_state(self.h), L.orc_amr_mesh(self.h)
        nb = self.nblocks
        dims = np.zeros(3, dtype=np.int32)
        L.orc_te_extents(m, 0, _ip(dims))
        shape = (nb, int(dims[2]), int(dims[1]), int(dims[0]))
        al = np.ctypeslib.as_array(L.orc_sparse_alloc(st), shape=(nb * 4,)).reshape(nb, 4).astype(bool)
        out = np.full((nb, 4) + shape[1:], np.nan)
        for f in range(4):
            u = np.ctypeslib.as_array(L.orc_sparse_U(st, f), shape=(int(np.prod(shape)),)).reshape(shape)
            out[:, f] = np.where(al[:, f, None, None, None], u, np.nan)
        return out


class SparseAdvection:
    """example/sparse_advection on a uniform mesh: four sparse fields"""

    NF = 4

    def __init__(self, mesh, speed=1.5, cfl=0.45, alloc_threshold=1e-5, dealloc_threshold=1e-6,
                 dealloc_count=5):
        self.mesh = mesh
        self.h = lib().orc_sparse_create(mesh.h, speed, cfl, alloc_threshold, dealloc_threshold,
                                         dealloc_count)

    def __del__(self):
        try:
            lib().orc_sparse_destroy(self.h)
        except Exception:
            pass

    def init(self):
        lib().orc_sparse_init(self.h)

    def step(self):
        lib().orc_sparse_step(self.h)

    @property
    def allocated(self):
        p = lib().orc_sparse_alloc(self.h)
        return np.ctypeslib.as_array(p, shape=(self.mesh.nblocks * 4,)).reshape(-1, 4).astype(bool)

    @property
    def U(self):
        """[nblocks][4][nk][nj][ni]; NaN where the field is not allocated (the convention of
        the reference dumps)"""
        shape = (self.mesh.nblocks,) + self.mesh.dims
        out = np.empty((self.mesh.nblocks, 4) + self.mesh.dims)
        a = self.allocated
        for f in range(4):
            u = np.ctypeslib.as_array(lib().orc_sparse_U(self.h, f),
                                      shape=(int(np.prod(shape)),)).reshape(shape)
            out[:, f] = np.where(a[:, f][:, None, None, None], u, np.nan)
        return out

    @property
    def dt(self):
        return lib().orc_sparse_dt(self.h)

    @property
    def time(self):
        return lib().orc_sparse_time(self.h)


def weno5z(q):
    ql, qr = C.c_double(), C.c_double()
    lib().orc_weno5z(*[float(x) for x in q], C.byref(ql), C.byref(qr))
    return ql.value, qr.value


def unpack_box_transformed(var, s, n, buf, dir_connection, dir_flip, ncell, fac):
    """SetBounds of ONE box of one block's array var[ncomp][nk][nj][ni] (modified in place)
    through a LogicalCoordinateTransformation (boundary_communication.cpp:282-308)"""
    assert var.dtype == np.float64 and var.flags.c_contiguous and var.ndim == 4
    ncomp, nk, nj, ni = var.shape
    i3 = C.c_int * 3
    b = np.ascontiguousarray(buf, dtype=np.float64)
    L = lib()
    L.orc_unpack_box_transformed.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, i3, i3,
                                             C.c_int, C.c_void_p, i3, i3, C.c_int, C.c_double]
    L.orc_unpack_box_transformed.restype = None
    L.orc_unpack_box_transformed(var.ctypes.data, ni, ni * nj, ni * nj * nk, i3(*s), i3(*n), ncomp,
                                 b.ctypes.data, i3(*dir_connection), i3(*[int(f) for f in dir_flip]),
                                 ncell, float(fac))
    return var
