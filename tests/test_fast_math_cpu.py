"""The product's FAST-math reconstruction (parthenon_b200/csrc/weno_fast.cuh), compiled as host
C++ by a tests-only harness, against the oracle's restatement of recon.hpp:27-99.

The device build differs from this host build only in the reciprocal seed (MUFU.RCP64H vs a
quotient cut to 20 mantissa bits); the cubic refinement step and everything else is the
same source.  The GPU tests check the device build itself within north_star's 1e-12."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DP = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("weno") / "weno_fast_host.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "cpu_harness", "weno_fast_host.cpp")])
    return C.CDLL(so)


def stencils():
    rng = np.random.default_rng(7)
    x = np.linspace(0, 1, 4000)[:, None] + np.arange(5)[None, :] * 1e-3
    step = np.ones((4000, 5))
    step[:, 3:] = 11.0
    step += 1e-3 * rng.normal(size=step.shape)
    return {
        "random": rng.normal(size=(4000, 5)),
        "smooth": np.sin(6 * x) + 2,
        "nearly_constant": 1 + 1e-9 * rng.normal(size=(4000, 5)),
        "step": step,                                  # the burgers scalar IC: 1 | 11
        "constant": np.ones((8, 5)),
        "large": rng.normal(size=(4000, 5)) * 1e20,
        "small": rng.normal(size=(4000, 5)) * 1e-6,
        "zeros": np.zeros((4, 5)),
    }


@pytest.mark.parametrize("kind", list(stencils()))
def test_weno5z_fast_vs_oracle(harness, kind):
    q = np.ascontiguousarray(stencils()[kind])
    n = len(q)
    ql, qr = np.zeros(n), np.zeros(n)
    harness.weno_fast_host(q.ctypes.data_as(DP), C.c_long(n), ql.ctypes.data_as(DP),
                           qr.ctypes.data_as(DP))
    ref = np.array([oracle.weno5z(r) for r in q])
    scale = np.maximum(np.abs(q).max(axis=1), 1e-300)
    assert np.all(np.isfinite(ql)) and np.all(np.isfinite(qr))
    # a few ulp of the stencil's magnitude; the contract on evolved fields is 1e-12
    assert (np.abs(ql - ref[:, 0]) / scale).max() < 5e-15
    assert (np.abs(qr - ref[:, 1]) / scale).max() < 5e-15


def test_linear_fast_bit_exact(harness):
    """the limiter's power-of-two rearrangement and sign-bit test change no bit"""
    rng = np.random.default_rng(3)
    q = rng.normal(size=(5000, 3))
    q[:50, 1] = q[:50, 0]          # zero left difference
    q[50:100, 2] = q[50:100, 1]    # zero right difference
    q = np.ascontiguousarray(q)
    ql, qr = np.zeros(len(q)), np.zeros(len(q))
    harness.linear_fast_host(q.ctypes.data_as(DP), C.c_long(len(q)), ql.ctypes.data_as(DP),
                             qr.ctypes.data_as(DP))
    L = oracle.lib()
    a, b = C.c_double(), C.c_double()
    for i in range(len(q)):
        L.orc_linear(*[float(x) for x in q[i]], C.byref(a), C.byref(b))
        assert ql[i] == a.value and qr[i] == b.value
