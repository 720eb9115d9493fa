"""ctypes binding of the C ABI in include/parthenon_b200.h (libpb200.so).

Plumbing only: structures mirror the header one to one.  There is no fallback — if the
shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpb200.so")

PB2_OK = 0
PB2_ERR_NO_DEVICE = -3
REGION_ALLOCATED, REGION_BUF_ALLOCATED, REGION_SAME_TO_SAME = 1, 2, 4
REGION_DST_UNALLOCATED = 8
RECON_WENO5, RECON_LINEAR = 0, 1
MATH_STRICT, MATH_FAST = 0, 1
PROLONG_MINMOD, PROLONG_LINEAR, PROLONG_PIECEWISE_CONSTANT = 0, 1, 2

c_double_p = C.POINTER(C.c_double)


class BndRegion(C.Structure):
    _fields_ = [("var", C.c_void_p), ("buf_off", C.c_int64), ("s", C.c_int32 * 3),
                ("n", C.c_int32 * 3), ("ncomp", C.c_int32), ("stride_j", C.c_int32),
                ("stride_k", C.c_int32), ("stride_c", C.c_int32), ("flag_slot", C.c_int32),
                ("status", C.c_uint32), ("value", C.c_double),
                ("lcoord_on", C.c_int32), ("lcoord_dir", C.c_int32 * 3),
                ("lcoord_flip", C.c_int32 * 3), ("lcoord_ncell", C.c_int32), ("fac", C.c_double)]


class CopyRegion(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("ss", C.c_int32 * 3),
                ("ds", C.c_int32 * 3), ("n", C.c_int32 * 3), ("ncomp", C.c_int32),
                ("src_stride_j", C.c_int32), ("src_stride_k", C.c_int32),
                ("src_stride_c", C.c_int32), ("dst_stride_j", C.c_int32),
                ("dst_stride_k", C.c_int32), ("dst_stride_c", C.c_int32),
                ("flag_slot", C.c_int32), ("status", C.c_uint32), ("threshold", C.c_double),
                ("default_value", C.c_double)]


class ProResRegion(C.Structure):
    _fields_ = [("fine", C.c_void_p), ("coarse", C.c_void_p), ("s", C.c_int32 * 3),
                ("n", C.c_int32 * 3), ("ncomp", C.c_int32), ("fine_stride_j", C.c_int32),
                ("fine_stride_k", C.c_int32), ("fine_stride_c", C.c_int32),
                ("coarse_stride_j", C.c_int32), ("coarse_stride_k", C.c_int32),
                ("coarse_stride_c", C.c_int32), ("fine_is", C.c_int32 * 3),
                ("coarse_is", C.c_int32 * 3), ("ndim", C.c_int32), ("status", C.c_uint32),
                ("fine_xmin", C.c_double * 3), ("fine_dx", C.c_double * 3),
                ("coarse_xmin", C.c_double * 3), ("coarse_dx", C.c_double * 3),
                ("ftop", C.c_int32 * 3), ("ctop", C.c_int32 * 3)]


class FlxCorRegion(C.Structure):
    _fields_ = [("fine", C.c_void_p), ("coarse", C.c_void_p), ("buf_off", C.c_int64),
                ("dir", C.c_int32), ("ndim", C.c_int32), ("fs", C.c_int32 * 3),
                ("ds", C.c_int32 * 3), ("n", C.c_int32 * 3), ("ncomp", C.c_int32),
                ("fine_stride_j", C.c_int32), ("fine_stride_k", C.c_int32),
                ("fine_stride_c", C.c_int32), ("coarse_stride_j", C.c_int32),
                ("coarse_stride_k", C.c_int32), ("coarse_stride_c", C.c_int32),
                ("status", C.c_uint32), ("area", C.c_double)]


class BcRegion(C.Structure):
    _fields_ = [("var", C.c_void_p), ("face", C.c_int32), ("type", C.c_int32),
                ("ncomp", C.c_int32), ("n", C.c_int32 * 3), ("is_", C.c_int32),
                ("ie", C.c_int32), ("stride_c", C.c_int32), ("flip_mask", C.c_uint32),
                ("stride_j", C.c_int32), ("stride_k", C.c_int32)]


class PackGeom(C.Structure):
    _fields_ = [("nblocks", C.c_int32), ("ncomp", C.c_int32), ("ndim", C.c_int32),
                ("nx", C.c_int32 * 3), ("ng", C.c_int32), ("block_stride", C.c_int64),
                ("dx", C.c_void_p), ("block_list", C.c_void_p), ("nlist", C.c_int32)]


class BurgersArgs(C.Structure):
    _fields_ = [("geom", PackGeom), ("recon", C.c_int32), ("math", C.c_int32),
                ("u", C.c_void_p), ("base", C.c_void_p), ("out", C.c_void_p),
                ("flux", C.c_void_p * 3), ("derived", C.c_void_p), ("dt_min", C.c_void_p),
                ("beta", C.c_double), ("dt", C.c_double),
                ("block_ids", C.c_void_p), ("num_block_ids", C.c_int32),
                ("nbr_direct", C.c_void_p), ("progress", C.c_void_p),
                ("progress_blocks", C.c_int32), ("sweeps", C.c_int32)]


_lib = None

# every symbol include/parthenon_b200.h declares
SYMBOLS = [
    "pb2_version", "pb2_last_error", "pb2_device_count", "pb2_set_device",
    "pb2_device_sm_count", "pb2_malloc", "pb2_free", "pb2_cache_trim", "pb2_host_alloc", "pb2_host_free",
    "pb2_memset", "pb2_memcpy_h2d", "pb2_memcpy_d2h", "pb2_memcpy_d2d", "pb2_stream_create", "pb2_stream_create_priority", "pb2_stream_wait_value",
    "pb2_stream_destroy", "pb2_stream_sync", "pb2_device_sync", "pb2_event_create",
    "pb2_event_destroy", "pb2_event_record", "pb2_event_sync", "pb2_event_query",
    "pb2_stream_wait_event", "pb2_event_elapsed_ms", "pb2_launch_count",
    "pb2_profile_enable", "pb2_profile_reset", "pb2_profile_kernels", "pb2_profile_get", "pb2_profile_get_work",
    "pb2_measure_fp64_peak",
    "pb2_bnd_table_create", "pb2_copy_table_create", "pb2_bnd_table_destroy",
    "pb2_bnd_table_elements", "pb2_pack", "pb2_unpack", "pb2_copy", "pb2_prores_table_create",
    "pb2_restrict", "pb2_prolongate", "pb2_restrict_te", "pb2_prolongate_te",
    "pb2_prolongate_internal", "pb2_prolongate_toth_roe", "pb2_flxcor_table_create", "pb2_flux_correct",
    "pb2_weighted_sum", "pb2_weighted_sum_ghosts", "pb2_weighted_sum_ghosts_blocks", "pb2_flux_divergence", "pb2_interior_scatter", "pb2_interior_gather",
    "pb2_halo_copy_uniform", "pb2_advection_fluxes", "pb2_copy_flags", "pb2_copy_select",
    "pb2_weighted_sum_blocks", "pb2_flux_divergence_blocks", "pb2_advection_fluxes_blocks",
    "pb2_block_quiet_flags", "pb2_block_minmax", "pb2_block_derivative", "pb2_bc_table_create", "pb2_apply_bcs",
    "pb2_burgers_calculate_fluxes", "pb2_burgers_update", "pb2_burgers_stage", "pb2_burgers_progress_target",
    "pb2_burgers_derived_dt", "pb2_burgers_history", "pb2_comm_unique_id", "pb2_comm_create", "pb2_comm_destroy",
    "pb2_comm_exchange", "pb2_comm_allreduce_min", "pb2_comm_allreduce_sum",
    "pb2_comm_barrier",
    "pb2_ipc_export", "pb2_ipc_open", "pb2_ipc_close", "pb2_peer_handshake", "pb2_copy_signal",
    "pb2_peer_signal", "pb2_peer_wait",
]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
                           "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    L.pb2_last_error.restype = C.c_char_p
    L.pb2_launch_count.restype = i64
    L.pb2_bnd_table_elements.restype = i64
    L.pb2_bnd_table_elements.argtypes = [vp]
    L.pb2_device_count.argtypes = [C.POINTER(C.c_int)]
    L.pb2_malloc.argtypes = [C.POINTER(vp), C.c_size_t]
    L.pb2_free.argtypes = [vp]
    L.pb2_memset.argtypes = [vp, C.c_int, C.c_size_t, vp]
    for f in ("pb2_memcpy_h2d", "pb2_memcpy_d2h", "pb2_memcpy_d2d"):
        getattr(L, f).argtypes = [vp, vp, C.c_size_t, vp]
    L.pb2_stream_sync.argtypes = [vp]
    L.pb2_peer_handshake.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int32, vp]
    L.pb2_copy_signal.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int32, vp]
    L.pb2_peer_signal.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int32, vp]
    L.pb2_peer_wait.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int32, vp]
    L.pb2_ipc_export.argtypes = [vp, vp]
    L.pb2_bnd_table_create.argtypes = [C.POINTER(vp), C.POINTER(BndRegion), i64]
    L.pb2_copy_table_create.argtypes = [C.POINTER(vp), C.POINTER(CopyRegion), i64]
    L.pb2_prores_table_create.argtypes = [C.POINTER(vp), C.POINTER(ProResRegion), i64]
    L.pb2_bc_table_create.argtypes = [C.POINTER(vp), C.POINTER(BcRegion), i64]
    L.pb2_apply_bcs.argtypes = [vp, vp]
    L.pb2_flxcor_table_create.argtypes = [C.POINTER(vp), C.POINTER(FlxCorRegion), i64]
    L.pb2_flux_correct.argtypes = [vp, vp, vp]
    L.pb2_bnd_table_destroy.argtypes = [vp]
    L.pb2_pack.argtypes = [vp, vp, vp, vp]
    L.pb2_unpack.argtypes = [vp, vp, vp, vp]
    L.pb2_copy.argtypes = [vp, vp, vp]
    L.pb2_copy_flags.argtypes = [vp, vp, vp]
    L.pb2_copy_select.argtypes = [vp, vp, vp]
    L.pb2_weighted_sum_blocks.argtypes = [C.POINTER(PackGeom), vp, vp, C.c_double, C.c_double,
                                          vp, vp, vp]
    L.pb2_flux_divergence_blocks.argtypes = [C.POINTER(PackGeom), C.POINTER(vp), vp, vp, vp]
    L.pb2_advection_fluxes_blocks.argtypes = [C.POINTER(PackGeom), vp, C.POINTER(vp),
                                              c_double_p, vp, vp]
    L.pb2_block_derivative.argtypes = [C.POINTER(PackGeom), vp, C.c_int, C.c_int, vp, vp]
    L.pb2_block_minmax.argtypes = [C.POINTER(PackGeom), vp, vp, vp, vp]
    L.pb2_block_quiet_flags.argtypes = [C.POINTER(PackGeom), vp, C.c_double, vp, vp, vp]
    L.pb2_halo_copy_uniform.argtypes = [vp, vp, vp, vp]
    L.pb2_restrict.argtypes = [vp, vp]
    L.pb2_prolongate.argtypes = [vp, C.c_int, vp]
    L.pb2_restrict_te.argtypes = [vp, vp]
    L.pb2_prolongate_te.argtypes = [vp, C.c_int, vp]
    L.pb2_prolongate_internal.argtypes = [vp, vp]
    L.pb2_prolongate_toth_roe.argtypes = [vp, vp]
    L.pb2_weighted_sum.argtypes = [vp, vp, C.c_double, C.c_double, vp, i64, vp]
    L.pb2_weighted_sum_ghosts.argtypes = [C.POINTER(PackGeom), vp, vp, C.c_double, C.c_double,
                                          vp, vp]
    L.pb2_advection_fluxes.argtypes = [C.POINTER(PackGeom), vp, C.POINTER(vp), c_double_p, vp]
    L.pb2_weighted_sum_ghosts_blocks.argtypes = [C.POINTER(PackGeom), vp, vp, C.c_double,
                                                 C.c_double, vp, vp, C.c_int32, vp]
    L.pb2_flux_divergence.argtypes = [C.POINTER(PackGeom), C.POINTER(vp), vp, vp]
    for f in ("pb2_burgers_calculate_fluxes", "pb2_burgers_update", "pb2_burgers_stage"):
        getattr(L, f).argtypes = [C.POINTER(BurgersArgs), vp]
    L.pb2_burgers_history.argtypes = [C.POINTER(PackGeom), vp, vp, c_double_p, c_double_p,
                                      c_double_p, vp]
    L.pb2_comm_unique_id.argtypes = [C.c_char_p]
    L.pb2_comm_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_char_p]
    L.pb2_comm_destroy.argtypes = [vp]
    L.pb2_comm_exchange.argtypes = [vp, vp, C.POINTER(i64), vp, C.POINTER(i64), vp]
    L.pb2_comm_allreduce_min.argtypes = [vp, vp, vp]
    L.pb2_comm_barrier.argtypes = [vp, vp]
    _lib = L
    return L


def check(rc):
    if rc != PB2_OK:
        raise RuntimeError(f"libpb200 error {rc}: {lib().pb2_last_error().decode()}")


def profile(enable=None, reset=False):
    """per-kernel device time measured with CUDA events: {name: (total_ms, launches)}"""
    L = lib()
    if reset:
        check(L.pb2_profile_reset())
    if enable is not None:
        check(L.pb2_profile_enable(int(enable)))
        return None
    out = {}
    for i in range(L.pb2_profile_kernels()):
        name, ms, n = C.c_char_p(), C.c_double(), C.c_int64()
        check(L.pb2_profile_get(i, C.byref(name), C.byref(ms), C.byref(n)))
        if n.value:
            out[name.value.decode()] = (ms.value, n.value)
    return out


def profile_work():
    """{name: work}: what the recorded launches processed (zones / values), see
    pb2_profile_get_work"""
    L = lib()
    out = {}
    for i in range(L.pb2_profile_kernels()):
        name, ms, n, w = C.c_char_p(), C.c_double(), C.c_int64(), C.c_double()
        check(L.pb2_profile_get(i, C.byref(name), C.byref(ms), C.byref(n)))
        check(L.pb2_profile_get_work(i, C.byref(w)))
        if n.value:
            out[name.value.decode()] = w.value
    return out


def fp64_peak_tflops():
    v = C.c_double()
    check(lib().pb2_measure_fp64_peak(C.byref(v)))
    return v.value


def launch_count():
    return int(lib().pb2_launch_count())


class Table:
    """Owning wrapper of a pb2_bnd_table."""

    def __init__(self, regions, kind):
        self.h = C.c_void_p()
        n = len(regions)
        if kind == "bnd":
            arr = (BndRegion * max(n, 1))(*regions)
            check(lib().pb2_bnd_table_create(C.byref(self.h), arr, n))
        elif kind == "copy":
            arr = (CopyRegion * max(n, 1))(*regions)
            check(lib().pb2_copy_table_create(C.byref(self.h), arr, n))
        elif kind == "prores":
            arr = (ProResRegion * max(n, 1))(*regions)
            check(lib().pb2_prores_table_create(C.byref(self.h), arr, n))
        elif kind == "bc":
            arr = (BcRegion * max(n, 1))(*regions)
            check(lib().pb2_bc_table_create(C.byref(self.h), arr, n))
        elif kind == "flxcor":
            arr = (FlxCorRegion * max(n, 1))(*regions)
            check(lib().pb2_flxcor_table_create(C.byref(self.h), arr, n))
        else:
            raise ValueError(kind)
        self.n = n

    @property
    def elements(self):
        return int(lib().pb2_bnd_table_elements(self.h))

    def __del__(self):
        try:
            if self.h:
                lib().pb2_bnd_table_destroy(self.h)
        except Exception:
            pass
