"""SparsePack / MakePackDescriptor: the host side resolves a descriptor against a
MeshData batch (parthenon_b200/host/pb2/sparse_pack.hpp), a downstream kernel written against
include/parthenon_b200_pack.h only indexes pack(b, n, k, j, i).  The checks follow the
reference's tst/unit/test_sparse_pack.cpp:47-330 (bounds of a block on which a variable is
not allocated, Contains, index maps, flattened packs, labels) on the sparse_advection
application's pool "sparse" (4 one-component sparse fields with fluxes) and on the burgers
application's 11-component U."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from parthenon_b200 import host
from tests.test_burgers_sim_gpu import burgers_overrides

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def user(tmp_path_factory):
    """the user kernels, compiled the way a downstream code would: nvcc + the pack header"""
    so = str(tmp_path_factory.mktemp("pack") / "pack_user_kernels.so")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2",
                           "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
                           "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cuda", "pack_user_kernels.cu"), "-o", so])
    L = C.CDLL(so)
    ip = C.POINTER(C.c_int)
    L.pack_check_var.argtypes = [host.SparsePackPOD, C.c_int, C.c_int, ip, ip]
    L.pack_check_flat.argtypes = [host.SparsePackPOD, ip]
    L.pack_flux_update.argtypes = [host.SparsePackPOD, C.c_double]
    return L


def pattern(shape, b0, v):
    """test_sparse_pack.cpp:170: n = i + 10 j + 100 k + 1e4 c + 1e5 v + 1e3 b"""
    nb, nc, nk, nj, ni = shape
    b, c, k, j, i = np.meshgrid(np.arange(nb), np.arange(nc), np.arange(nk), np.arange(nj),
                                np.arange(ni), indexing="ij")
    return (i + 1e1 * j + 1e2 * k + 1e4 * c + 1e5 * v + 1e3 * (b + b0)).astype(np.float64)


def sparse_sim():
    ov = {"parthenon/mesh/nx1": 16, "parthenon/mesh/nx2": 16, "parthenon/meshblock/nx1": 8,
          "parthenon/meshblock/nx2": 8, "parthenon/mesh/refinement": "none"}
    return host.Simulation(app="sparse_advection", overrides=ov)


def test_bounds_contains_and_labels_follow_allocation(user):
    sim = sparse_sim()
    try:
        nb = sim.info()["nblocks"]
        assert nb == 4
        fields = [f"sparse_{i}" for i in range(4)]
        for f in fields:
            for b in range(nb):
                sim.set_sparse_allocation(f, b, True)
        for v, f in enumerate(fields):
            sim.set_field("base", f, pattern(sim.field_shape("base", f), 0, v))
        # deallocate one variable on an arbitrary block (test_sparse_pack.cpp:184)
        sim.set_sparse_allocation("sparse_1", 2, False)
        assert not sim.allocation("base", "sparse_1")[2]

        # selecting the pool by its base name: one group, members ordered by sparse id
        pod, bounds = sim.sparse_pack("base", ["sparse"])
        assert (pod.nblocks, pod.nvar, pod.size) == (nb, 1, 4 * nb - 1)
        assert pod.maxvars == 4
        assert bounds[0, :, 0].tolist() == [0, 0, 0, 0]
        assert bounds[1, :, 0].tolist() == [3, 3, 2, 3]  # block 2 lost one component
        assert sim.sparse_pack_label("base", ["sparse"], 2, 1) == "sparse_2"
        assert sim.sparse_pack_label("base", ["sparse"], 1, 1) == "sparse_1"

        # one group per field, descriptor order (3, 1): block 2 does not contain sparse_1
        names = ["sparse_3", "sparse_1"]
        pod, bounds = sim.sparse_pack("base", names, flags=["WithFluxes"])
        assert bounds[0, 2].tolist() == [0, -1, 0] and bounds[1, 2].tolist() == [0, -2, 0]
        assert bounds[0, 1].tolist() == [0, 1, 0] and bounds[1, 1].tolist() == [0, 1, 1]
        # ... and a user kernel reads every allocated component through both accessors
        for var, vid in ((0, 3), (1, 1)):
            nwrong, nseen = C.c_int(-1), C.c_int(-1)
            assert user.pack_check_var(pod, var, vid, C.byref(nwrong), C.byref(nseen)) == 0
            assert nwrong.value == 0
            ncell = pod.ni * pod.nj * pod.nk
            assert nseen.value == ncell * (nb if var == 0 else nb - 1)
        # a regular expression selects the same fields
        pod_re, b_re = sim.sparse_pack("base", ["re:sparse_[13]"])
        assert b_re[1, :, 0].tolist() == [1, 1, 0, 1]

        # flattened: one unified index space, block 2 one entry short
        podf, bf = sim.sparse_pack("base", ["sparse"], flatten=True)
        assert podf.nblocks == 1 and podf.maxvars == 4 * nb - 1 and podf.size == 4 * nb - 1
        assert bf[1, :, 1].tolist() == [3, 7, 10, 14]  # inclusive upper index per block
        nwrong = C.c_int(-1)
        assert user.pack_check_flat(podf, C.byref(nwrong)) == 0 and nwrong.value == 0

        # allocation changes -> the cached pack is rebuilt, not reused
        sim.set_sparse_allocation("sparse_1", 2, True)
        pod2, b2 = sim.sparse_pack("base", ["sparse"])
        assert b2[1, :, 0].tolist() == [3, 3, 3, 3] and pod2.size == 4 * nb
    finally:
        sim.close()


def test_user_kernel_updates_through_pack_and_flux_accessors(user):
    """a task-shaped user kernel (u -= dt/dx (F(i+1) - F(i)) through pack(b,n,k,j,i) /
    pack.flux(b,1,n,k,j,i) / GetCoordinates) on the burgers application's 11-component U"""
    sim = host.Simulation(overrides=burgers_overrides(8, 2, 4, 8, "weno5", "strict", True))
    try:
        sim.pre_execute()
        U = sim.get_field("base", "U")
        rng = np.random.default_rng(3)
        F = rng.standard_normal(U.shape)
        sim.set_field("base", "U", F, which=host.FIELD_FLUX1)
        pod, bounds = sim.sparse_pack("base", ["U"], with_fluxes=True)
        assert (pod.nblocks, pod.maxvars, pod.size) == (8, 11, 88)
        assert (pod.ni, pod.nj, pod.nk) == (16, 16, 16) and (pod.is_, pod.ie) == (4, 11)
        assert bounds[1, :, 0].tolist() == [10] * 8
        dt = 0.125
        assert user.pack_flux_update(pod, dt) == 0
        got = sim.get_field("base", "U")
        dx = 1.0 / 16
        want = U.copy()
        g = 4
        want[..., g:-g, g:-g, g:-g] -= dt / dx * (F[..., g:-g, g:-g, g + 1:-g + 1] -
                                                  F[..., g:-g, g:-g, g:-g])
        assert np.array_equal(got, want)
    finally:
        sim.close()
