"""ctypes binding of libpb200_host.so (include/parthenon_b200_host.h): the C++ host framework
(ParameterInput / StateDescriptor / Mesh / MeshData / TaskList / BurgersDriver mirroring
Parthenon's API) driven from Python for bench.py and the tests.

Plumbing only — there is no Python fallback for anything: a missing library or a failing call
raises RuntimeError.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpb200_host.so")

SYMBOLS = [
    "pb2h_last_error", "pb2h_sim_create", "pb2h_topology_create", "pb2h_topology_regrid",
    "pb2h_topology_derefine_counts",
    "pb2h_sim_destroy", "pb2h_sim_tag_and_remesh",
    "pb2h_sim_pre_execute", "pb2h_sim_cycle", "pb2h_sim_cycle_phase", "pb2h_sim_execute", "pb2h_sim_sync",
    "pb2h_sim_stream", "pb2h_sim_time", "pb2h_sim_dt", "pb2h_sim_ncycle", "pb2h_sim_set_dt",
    "pb2h_sim_zone_cycles_per_second", "pb2h_sim_info", "pb2h_sim_block", "pb2h_sim_neighbor", "pb2h_sim_block_bcs",
    "pb2h_topology_create_forest", "pb2h_sim_exchange_mode", "pb2h_sim_flux_correction", "pb2h_sim_edge_flux_plan",
    "pb2h_sim_calc_indices", "pb2h_sim_ranklist", "pb2h_sim_plan", "pb2h_sim_plan_boxes",
    "pb2h_sim_field_ptr", "pb2h_sim_field_dims",
    "pb2h_sim_get_field", "pb2h_sim_set_field", "pb2h_sim_allocation", "pb2h_sim_exchange", "pb2h_sim_exchange_phase",
    "pb2h_sim_exchange_elements", "pb2h_sim_history", "pb2h_sim_upload_interior",
    "pb2h_sim_download_interior", "pb2h_sim_prefetch_interior", "pb2h_sim_commit_interior",
    "pb2h_sim_writeback_interior", "pb2h_sim_lane_sync",
    "pb2h_sim_sparse_pack", "pb2h_sim_sparse_pack_label", "pb2h_sim_set_sparse_allocation",
]

BURGERS_DECK = """
<parthenon/job>
problem_id = burgers
<parthenon/mesh>
nghost = 4
refinement = none
numlevel = 1
nx1 = 64
x1min = -0.5
x1max = 0.5
ix1_bc = periodic
ox1_bc = periodic
nx2 = 64
x2min = -0.5
x2max = 0.5
ix2_bc = periodic
ox2_bc = periodic
nx3 = 64
x3min = -0.5
x3max = 0.5
ix3_bc = periodic
ox3_bc = periodic
<parthenon/meshblock>
nx1 = 32
nx2 = 32
nx3 = 32
<parthenon/time>
nlim = -1
tlim = 1e9
integrator = rk2
ncycle_out = 0
perf_cycle_offset = 0
<parthenon/refinement0>
method = derivative_order_1
field = U
vector_i = 3
refine_tol = 0.5
derefine_tol = 0.2
<burgers>
cfl = 0.8
recon = weno5
num_scalars = 8
"""

# the field-only test application (face / edge / node fields): the burgers deck's mesh and time
# blocks without its <parthenon/refinement0> criterion — that criterion names the field U, which
# this application does not have, and a criterion on a missing field votes "same"
# (amr_criteria.cpp:89-91), i.e. it would veto every derefinement
TECOMM_DECK = BURGERS_DECK[:BURGERS_DECK.index("<parthenon/refinement0>")] + \
    BURGERS_DECK[BURGERS_DECK.index("<burgers>"):]

# the forest application (host/forest): example/boundary_exchange's deck (the mesh itself is a
# forest built in code); the outer edges take the reference's default condition, outflow
FOREST_DECK = """
<parthenon/job>
problem_id = forest
<parthenon/mesh>
refinement = static
numlevel = 1
nghost = 2
<parthenon/meshblock>
nx1 = 4
nx2 = 4
nx3 = 1
<forest>
variant = 0
"""

# example/advection: the values of the reference's parthinput.advection that matter on this
# path (outputs and the derived demo fields are off)
ADVECTION_DECK = """
<parthenon/job>
problem_id = advection
<parthenon/mesh>
nghost = 2
refinement = none
numlevel = 1
nx1 = 64
x1min = -0.5
x1max = 0.5
ix1_bc = periodic
ox1_bc = periodic
nx2 = 64
x2min = -0.5
x2max = 0.5
ix2_bc = periodic
ox2_bc = periodic
nx3 = 1
x3min = -0.5
x3max = 0.5
ix3_bc = periodic
ox3_bc = periodic
<parthenon/meshblock>
nx1 = 16
nx2 = 16
nx3 = 1
<parthenon/time>
nlim = -1
tlim = 1e9
integrator = rk2
ncycle_out = 0
perf_cycle_offset = 0
<Advection>
cfl = 0.45
vx = 1.0
vy = 1.0
vz = 1.0
profile = hard_sphere
refine_tol = 0.3
derefine_tol = 0.03
num_vars = 1
vec_size = 1
fill_derived = false
"""

# example/sparse_advection: the reference's parthinput.sparse_advection on a uniform mesh
SPARSE_ADVECTION_DECK = """
<parthenon/job>
problem_id = sparse
<parthenon/sparse>
enable_sparse = true
alloc_threshold = 1e-5
dealloc_threshold = 1e-6
dealloc_count = 5
<parthenon/mesh>
nghost = 2
refinement = none
numlevel = 1
nx1 = 64
x1min = -1.0
x1max = 1.0
ix1_bc = periodic
ox1_bc = periodic
nx2 = 64
x2min = -1.0
x2max = 1.0
ix2_bc = periodic
ox2_bc = periodic
nx3 = 1
x3min = -1.0
x3max = 1.0
ix3_bc = periodic
ox3_bc = periodic
<parthenon/meshblock>
nx1 = 8
nx2 = 8
nx3 = 1
<parthenon/time>
nlim = -1
tlim = 1e9
integrator = rk2
ncycle_out = 0
perf_cycle_offset = 0
<sparse_advection>
restart_test = false
cfl = 0.45
speed = 1.5
refine_tol = 0.3
derefine_tol = 0.03
"""

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
                           "(there is no Python/CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i64 = C.c_void_p, C.c_int64
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    L.pb2h_last_error.restype = C.c_char_p
    L.pb2h_sim_create.argtypes = [C.POINTER(vp), C.c_char_p, C.c_char_p, C.c_char_p, C.c_int,
                                  C.c_int, C.c_char_p, ip, C.c_int]
    L.pb2h_topology_create.argtypes = [C.POINTER(vp), C.c_char_p, C.c_char_p, C.c_int, C.c_int,
                                       ip, C.c_int]
    for f in ("pb2h_sim_destroy", "pb2h_sim_pre_execute", "pb2h_sim_execute", "pb2h_sim_sync"):
        getattr(L, f).argtypes = [vp]
    L.pb2h_topology_regrid.argtypes = [vp, ip, C.c_int, ip]
    L.pb2h_sim_tag_and_remesh.argtypes = [vp, C.c_int, ip]
    L.pb2h_topology_derefine_counts.argtypes = [vp, ip, C.c_int, C.c_int]
    L.pb2h_sim_cycle.argtypes = [vp, C.c_int]
    L.pb2h_sim_cycle_phase.argtypes = [vp, C.c_int]
    L.pb2h_sim_stream.restype = vp
    L.pb2h_sim_stream.argtypes = [vp]
    for f in ("pb2h_sim_time", "pb2h_sim_dt", "pb2h_sim_zone_cycles_per_second"):
        getattr(L, f).restype = C.c_double
        getattr(L, f).argtypes = [vp]
    L.pb2h_sim_ncycle.argtypes = [vp]
    L.pb2h_sim_set_dt.argtypes = [vp, C.c_double]
    L.pb2h_sim_info.argtypes = [vp, ip]
    L.pb2h_sim_block.argtypes = [vp, C.c_int, ip, dp, dp, ip, ip]
    L.pb2h_sim_block_bcs.argtypes = [vp, C.c_int, ip]
    L.pb2h_topology_create_forest.argtypes = [C.POINTER(vp), C.c_char_p, C.c_char_p, C.c_int,
                                              C.c_int, C.c_int]
    L.pb2h_sim_neighbor.argtypes = [vp, C.c_int, C.c_int, ip]
    L.pb2h_sim_calc_indices.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, ip, ip]
    L.pb2h_sim_ranklist.argtypes = [vp, ip, C.c_int]
    L.pb2h_sim_plan.restype = i64
    L.pb2h_sim_plan.argtypes = [vp, C.c_int, C.c_int, C.POINTER(i64), i64, C.POINTER(i64)]
    L.pb2h_sim_plan_boxes.restype = i64
    L.pb2h_sim_plan_boxes.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(i64), i64]
    L.pb2h_sim_field_dims.argtypes = [vp, C.c_char_p, C.c_char_p, C.POINTER(C.c_int)]
    L.pb2h_sim_field_ptr.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_int, C.POINTER(vp),
                                     C.POINTER(i64)]
    L.pb2h_sim_get_field.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_int, vp, i64]
    L.pb2h_sim_set_field.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_int, vp, i64]
    L.pb2h_sim_allocation.argtypes = [vp, C.c_char_p, C.c_char_p, ip, C.c_int]
    L.pb2h_sim_upload_interior.argtypes = [vp, C.c_char_p, C.c_char_p, vp, i64]
    L.pb2h_sim_download_interior.argtypes = [vp, C.c_char_p, C.c_char_p, vp, i64]
    L.pb2h_sim_prefetch_interior.argtypes = [vp, C.c_char_p, C.c_char_p, vp, i64, C.c_int]
    L.pb2h_sim_commit_interior.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_int]
    L.pb2h_sim_writeback_interior.argtypes = [vp, C.c_char_p, C.c_char_p, vp, i64, C.c_int]
    L.pb2h_sim_lane_sync.argtypes = [vp, C.c_int]
    L.pb2h_sim_exchange.argtypes = [vp, C.c_char_p, C.c_int]
    L.pb2h_sim_exchange_phase.argtypes = [vp, C.c_char_p, C.c_int]
    L.pb2h_sim_exchange_mode.argtypes = [vp, C.c_char_p]
    L.pb2h_sim_flux_correction.argtypes = [vp, C.c_char_p]
    L.pb2h_sim_edge_flux_plan.restype = i64
    L.pb2h_sim_edge_flux_plan.argtypes = [vp, C.c_int, C.POINTER(i64), i64]
    L.pb2h_sim_exchange_elements.restype = i64
    L.pb2h_sim_exchange_elements.argtypes = [vp, C.c_char_p, C.POINTER(i64), C.POINTER(i64)]
    L.pb2h_sim_history.argtypes = [vp, dp]
    L.pb2h_sim_sparse_pack.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, vp,
                                       C.POINTER(C.c_int32), i64]
    L.pb2h_sim_sparse_pack_label.restype = C.c_char_p
    L.pb2h_sim_sparse_pack_label.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int,
                                             C.c_int, C.c_int]
    L.pb2h_sim_set_sparse_allocation.argtypes = [vp, C.c_char_p, C.c_int, C.c_int]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise RuntimeError("libpb200_host: " + lib().pb2h_last_error().decode())


def _overrides(ov):
    if ov is None:
        return b""
    if isinstance(ov, dict):
        ov = [f"{k}={v}" for k, v in ov.items()]
    return "\n".join(ov).encode()


def _leaves(leaves):
    if leaves is None:
        return None, 0
    a = np.ascontiguousarray(leaves, dtype=np.int32)
    return a, a.shape[0]


FIELD_DATA, FIELD_FLUX1, FIELD_FLUX2, FIELD_FLUX3, FIELD_COARSE = 0, 1, 2, 3, 4


class SparsePackPOD(C.Structure):
    """pb2_sparse_pack of include/parthenon_b200_pack.h"""
    _fields_ = [("ptr", C.c_void_p), ("bounds", C.c_void_p), ("coords", C.c_void_p),
                ("nblocks", C.c_int32), ("nblocks_md", C.c_int32), ("maxvars", C.c_int32),
                ("nvar", C.c_int32), ("size", C.c_int32), ("flat", C.c_int32),
                ("with_fluxes", C.c_int32), ("coarse", C.c_int32),
                ("ni", C.c_int32), ("nj", C.c_int32), ("nk", C.c_int32),
                ("is_", C.c_int32), ("ie", C.c_int32), ("js", C.c_int32), ("je", C.c_int32),
                ("ks", C.c_int32), ("ke", C.c_int32)]


class _Base:
    h = None

    def info(self):
        o = (C.c_int * 12)()
        check(lib().pb2h_sim_info(self.h, o))
        keys = ["ndim", "nbtotal", "nblocks", "ni", "nj", "nk", "cni", "cnj", "cnk",
                "multilevel", "first_gid", "nghost"]
        return dict(zip(keys, [int(x) for x in o]))

    def block(self, lid):
        loc = (C.c_int * 4)()
        lo, hi = (C.c_double * 3)(), (C.c_double * 3)()
        gid, nn = C.c_int(), C.c_int()
        check(lib().pb2h_sim_block(self.h, lid, loc, lo, hi, C.byref(gid), C.byref(nn)))
        return dict(loc=tuple(loc), xmin=np.array(lo), xmax=np.array(hi), gid=gid.value,
                    nneighbors=nn.value)

    def block_bcs(self, lid):
        """MeshBlock::boundary_flag per face: -1 block, 1 reflect, 2 outflow, 3 periodic, 4 user"""
        o = (C.c_int * 6)()
        check(lib().pb2h_sim_block_bcs(self.h, lid, o))
        return tuple(int(x) for x in o)

    def neighbors(self, lid):
        out = []
        o = (C.c_int * 6)()
        for n in range(self.block(lid)["nneighbors"]):
            check(lib().pb2h_sim_neighbor(self.h, lid, n, o))
            out.append(tuple(int(x) for x in o))
        return out

    def calc_indices(self, lid, n, ir_type, prores=False):
        s, e = (C.c_int * 3)(), (C.c_int * 3)()
        check(lib().pb2h_sim_calc_indices(self.h, lid, n, ir_type, int(prores), s, e))
        return tuple(s), tuple(e)

    def ranklist(self):
        n = self.info()["nbtotal"]
        a = np.zeros(n, dtype=np.int32)
        check(lib().pb2h_sim_ranklist(self.h, a.ctypes.data_as(C.POINTER(C.c_int)), n))
        return a

    def plan(self, ncomp, kind):
        """kind: 'local' | 'send' | 'recv' -> (rows[n,7], seg_off)"""
        k = {"local": 0, "send": 1, "recv": 2}[kind]
        n = lib().pb2h_sim_plan(self.h, ncomp, k, None, 0, None)
        if n < 0:
            check(-1)
        rows = np.zeros((max(n, 1), 7), dtype=np.int64)
        seg = np.zeros(4096, dtype=np.int64)
        p64 = C.POINTER(C.c_int64)
        lib().pb2h_sim_plan(self.h, ncomp, k, rows.ctypes.data_as(p64), n, seg.ctypes.data_as(p64))
        return rows[:n], seg

    def plan_boxes(self, ncomp, tt, kind):
        """channels (pieces) of one field of topological type tt (0 cell, 1 face, 2 edge, 3 node)
        with their index boxes: rows[n, 18] = [sender_gid, receiver_gid, offset_index, piece,
        comp0, ncomp, send_s(i,j,k), recv_s(i,j,k), n(i,j,k), slab_off, peer, coarse flags]"""
        k = {"local": 0, "send": 1, "recv": 2}[kind]
        n = lib().pb2h_sim_plan_boxes(self.h, ncomp, tt, k, None, 0)
        if n < 0:
            check(-1)
        rows = np.zeros((max(n, 1), 18), dtype=np.int64)
        lib().pb2h_sim_plan_boxes(self.h, ncomp, tt, k, rows.ctypes.data_as(C.POINTER(C.c_int64)), n)
        return rows[:n]

    def edge_flux_plan(self, kind):
        """flux correction of a face field as topology: kind 'restrict' | 'deliver' (same-device
        neighbours) | 'send_restrict' | 'send' | 'recv' (neighbours on another device)
        -> rows[n, 16] (include/parthenon_b200_host.h)"""
        k = {"restrict": 0, "deliver": 1, "send_restrict": 2, "send": 3, "recv": 4}[kind]
        n = lib().pb2h_sim_edge_flux_plan(self.h, k, None, 0)
        if n < 0:
            check(-1)
        rows = np.zeros((max(n, 1), 16), dtype=np.int64)
        lib().pb2h_sim_edge_flux_plan(self.h, k, rows.ctypes.data_as(C.POINTER(C.c_int64)), n)
        return rows[:n]

    def close(self):
        if self.h:
            lib().pb2h_sim_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Topology(_Base):
    """Mesh topology only; touches no device (CPU tests)."""

    def __init__(self, deck=BURGERS_DECK, overrides=None, rank=0, nranks=1, leaves=None):
        self.h = C.c_void_p()
        la, n = _leaves(leaves)
        check(lib().pb2h_topology_create(C.byref(self.h), deck.encode(), _overrides(overrides),
                                         rank, nranks,
                                         la.ctypes.data_as(C.POINTER(C.c_int)) if n else None, n))

    def regrid(self, tags):
        """apply one AmrTag per block (SetRefinement rules), update the tree, rebuild the block
        list; returns True if the mesh changed"""
        t = np.ascontiguousarray(tags, dtype=np.int32)
        changed = C.c_int(0)
        check(lib().pb2h_topology_regrid(self.h, t.ctypes.data_as(C.POINTER(C.c_int)), len(t),
                                         C.byref(changed)))
        return bool(changed.value)

    @property
    def derefine_counts(self):
        out = np.zeros(self.info()["nblocks"], dtype=np.int32)
        check(lib().pb2h_topology_derefine_counts(self.h, out.ctypes.data_as(C.POINTER(C.c_int)),
                                                  len(out), 0))
        return out

    @derefine_counts.setter
    def derefine_counts(self, counts):
        c = np.ascontiguousarray(counts, dtype=np.int32)
        check(lib().pb2h_topology_derefine_counts(self.h, c.ctypes.data_as(C.POINTER(C.c_int)),
                                                  len(c), 1))


class ForestTopology(_Base):
    """Topology of the forest application (2 x 2 forests of example/boundary_exchange, variant
    0..3); touches no device (CPU tests)."""

    def __init__(self, variant, overrides=None, rank=0, nranks=1):
        self.h = C.c_void_p()
        check(lib().pb2h_topology_create_forest(C.byref(self.h), FOREST_DECK.encode(),
                                                _overrides(overrides), variant, rank, nranks))


class Simulation(_Base):
    """A running application on the GPU (ParthenonManager + BurgersDriver)."""

    def __init__(self, app="burgers", deck=None, overrides=None, rank=0, nranks=1,
                 nccl_id=None, leaves=None):
        if deck is None:
            deck = {"advection": ADVECTION_DECK,
                    "sparse_advection": SPARSE_ADVECTION_DECK,
                    "tecomm": TECOMM_DECK, "forest": FOREST_DECK}.get(app, BURGERS_DECK)
        self.h = C.c_void_p()
        la, n = _leaves(leaves)
        check(lib().pb2h_sim_create(C.byref(self.h), app.encode(), deck.encode(),
                                    _overrides(overrides), rank, nranks, nccl_id,
                                    la.ctypes.data_as(C.POINTER(C.c_int)) if n else None, n))

    def pre_execute(self):
        check(lib().pb2h_sim_pre_execute(self.h))

    def cycle(self, n=1):
        check(lib().pb2h_sim_cycle(self.h, n))

    def step(self):
        """first half of a cycle: Step + time advance (adaptive meshes: before the remesh)"""
        check(lib().pb2h_sim_cycle_phase(self.h, 0))

    def regrid(self):
        """second half: LoadBalancingAndAdaptiveMeshRefinement + SetGlobalTimeStep"""
        check(lib().pb2h_sim_cycle_phase(self.h, 1))

    def tag_and_remesh(self, cycle):
        """application tecomm, adaptive mesh: tag with the criterion of `cycle` and remesh"""
        changed = C.c_int(0)
        check(lib().pb2h_sim_tag_and_remesh(self.h, cycle, C.byref(changed)))
        return bool(changed.value)

    def execute(self):
        check(lib().pb2h_sim_execute(self.h))

    def sync(self):
        check(lib().pb2h_sim_sync(self.h))

    @property
    def stream(self):
        return lib().pb2h_sim_stream(self.h)

    @property
    def time(self):
        return lib().pb2h_sim_time(self.h)

    @property
    def dt(self):
        return lib().pb2h_sim_dt(self.h)

    def set_dt(self, dt):
        check(lib().pb2h_sim_set_dt(self.h, float(dt)))

    @property
    def ncycle(self):
        return lib().pb2h_sim_ncycle(self.h)

    def field_shape(self, container, field, which=FIELD_DATA):
        i = self.info()
        _, n = self.field_ptr(container, field, which)
        if which == FIELD_COARSE:
            cell = (i["cnk"], i["cnj"], i["cni"])
        else:
            d = (C.c_int * 6)()
            check(lib().pb2h_sim_field_dims(self.h, container.encode(), field.encode(), d))
            cell = (d[2], d[3], d[4])
        ncomp = n // (i["nblocks"] * cell[0] * cell[1] * cell[2])
        return (i["nblocks"], ncomp) + cell

    def field_ptr(self, container, field, which=FIELD_DATA):
        p, n = C.c_void_p(), C.c_int64()
        check(lib().pb2h_sim_field_ptr(self.h, container.encode(), field.encode(), which,
                                       C.byref(p), C.byref(n)))
        return p.value, n.value

    def allocation(self, container, field):
        """bool [nblocks]: where a (sparse) field is allocated on this rank"""
        nb = self.info()["nblocks"]
        out = np.zeros(nb, dtype=np.int32)
        check(lib().pb2h_sim_allocation(self.h, container.encode(), field.encode(),
                                        out.ctypes.data_as(C.POINTER(C.c_int)), nb))
        return out.astype(bool)

    def get_field(self, container, field, which=FIELD_DATA, out=None):
        shape = self.field_shape(container, field, which)
        if out is None:
            out = np.empty(shape)
        check(lib().pb2h_sim_get_field(self.h, container.encode(), field.encode(), which,
                                       out.ctypes.data, out.size))
        return out

    def set_field(self, container, field, array, which=FIELD_DATA):
        a = np.ascontiguousarray(array, dtype=np.float64)
        check(lib().pb2h_sim_set_field(self.h, container.encode(), field.encode(), which,
                                       a.ctypes.data, a.size))

    def sparse_pack(self, container, names, flags=(), with_fluxes=False, coarse=False,
                    flatten=False):
        """parthenon::MakePackDescriptor(...).GetPack(container) -> (POD a kernel takes by value,
        host bounds [2][nblocks][nvar+1]).  names: variable / sparse-pool names ("re:<regex>" for a
        regular expression), flags: Metadata flag names"""
        pod = SparsePackPOD()
        opt = int(with_fluxes) | 2 * int(coarse) | 4 * int(flatten)
        nb = self.info()["nblocks"]
        bounds = np.zeros(2 * max(nb, 1) * (len(names) + 1), dtype=np.int32)
        check(lib().pb2h_sim_sparse_pack(self.h, container.encode(), "\n".join(names).encode(),
                                         "\n".join(flags).encode(), opt, C.byref(pod),
                                         bounds.ctypes.data_as(C.POINTER(C.c_int32)), bounds.size))
        n = 2 * pod.nblocks_md * (pod.nvar + 1)
        return pod, bounds[:n].reshape(2, pod.nblocks_md, pod.nvar + 1)

    def sparse_pack_label(self, container, names, b, idx, flags=(), with_fluxes=False,
                          flatten=False):
        opt = int(with_fluxes) | 4 * int(flatten)
        lib().pb2h_sim_sparse_pack_label.restype = C.c_char_p
        return lib().pb2h_sim_sparse_pack_label(self.h, container.encode(),
                                                "\n".join(names).encode(),
                                                "\n".join(flags).encode(), opt, b, idx).decode()

    def set_sparse_allocation(self, field, lid, allocated):
        check(lib().pb2h_sim_set_sparse_allocation(self.h, field.encode(), lid, int(allocated)))

    def interior_size(self, container, field):
        shape = self.field_shape(container, field)
        g = self.info()["nghost"]
        nd = self.info()["ndim"]
        cells = 1
        for d, n in enumerate(reversed(shape[2:])):
            cells *= (n - 2 * g) if d < nd else n
        return shape[0] * shape[1] * cells

    def upload_interior(self, container, field, host_ptr, nreal):
        """host_ptr: address of nreal doubles [block][comp][nx3][nx2][nx1]; asynchronous"""
        check(lib().pb2h_sim_upload_interior(self.h, container.encode(), field.encode(),
                                             host_ptr, nreal))

    def download_interior(self, container, field, host_ptr, nreal):
        check(lib().pb2h_sim_download_interior(self.h, container.encode(), field.encode(),
                                               host_ptr, nreal))

    # pipelined lanes (independent batches of state; see include/parthenon_b200_host.h)
    def prefetch_interior(self, container, field, host_ptr, nreal, lane):
        check(lib().pb2h_sim_prefetch_interior(self.h, container.encode(), field.encode(),
                                               host_ptr, nreal, lane))

    def commit_interior(self, container, field, lane):
        check(lib().pb2h_sim_commit_interior(self.h, container.encode(), field.encode(), lane))

    def writeback_interior(self, container, field, host_ptr, nreal, lane):
        check(lib().pb2h_sim_writeback_interior(self.h, container.encode(), field.encode(),
                                                host_ptr, nreal, lane))

    def lane_sync(self, lane):
        check(lib().pb2h_sim_lane_sync(self.h, lane))

    def exchange(self, container="base", prolongate=True):
        check(lib().pb2h_sim_exchange(self.h, container.encode(), int(prolongate)))

    def exchange_phase(self, container, phase):
        check(lib().pb2h_sim_exchange_phase(self.h, container.encode(), phase))

    def exchange_elements(self, container="base"):
        lo, nl = C.c_int64(), C.c_int64()
        t = lib().pb2h_sim_exchange_elements(self.h, container.encode(), C.byref(lo), C.byref(nl))
        if t < 0:
            check(-1)
        return lo.value, nl.value

    def flux_correction(self, container="base"):
        """AddFluxCorrectionTasks on `container` (face fluxes of cell-centred fields, edge-centred
        fluxes of face fields)"""
        check(lib().pb2h_sim_flux_correction(self.h, container.encode()))

    def exchange_mode(self, container="base"):
        """how the inter-device halo travels (include/parthenon_b200_host.h)"""
        m = lib().pb2h_sim_exchange_mode(self.h, container.encode())
        if m < 0:
            check(-1)
        return {0: "none", 1: "slabs + grouped ncclSend/ncclRecv",
                2: "peer push: pack, copy-engine copies into the peers' receive slabs over "
                   "NVLink (CUDA IPC), arrival flags, local unpack",
                3: "peer push: pack kernel stores into the peers' receive slabs, arrival flags",
                4: "peer push: copy kernel stores into the peers' ghost cells, arrival flags"}[m]

    def history(self):
        o = np.zeros(8)
        check(lib().pb2h_sim_history(self.h, o.ctypes.data_as(C.POINTER(C.c_double))))
        return o

    def zone_cycles_per_second(self):
        return lib().pb2h_sim_zone_cycles_per_second(self.h)
