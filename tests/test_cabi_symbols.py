"""CPU test of the drop-in boundary: both shared libraries load without a GPU and export every
function `include/*.h` declares; the ctypes bindings know every one of them; and a compute entry
called without a device fails loudly instead of falling back to anything."""
import ctypes as C
import os
import re

import pytest

from parthenon_b200 import capi, host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECL = re.compile(r"^[A-Za-z_][\w\s\*]*?\b(pb2h?_\w+)\s*\(", re.M)


def declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)  # comments mention function names too
    names = [n for n in DECL.findall(text) if not n.endswith("_t")]
    assert len(names) > 20, header
    return sorted(set(names))


@pytest.mark.parametrize("header,libpath,bound", [
    ("parthenon_b200.h", capi.LIB_PATH, capi.SYMBOLS),
    ("parthenon_b200_host.h", host.LIB_PATH, host.SYMBOLS)])
def test_library_exports_every_declared_symbol(header, libpath, bound):
    assert os.path.exists(libpath), f"{libpath} is not built (python __graft_entry__.py)"
    lib = C.CDLL(libpath)
    names = declared(header)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/{header} but not exported: {missing}"
    unbound = [n for n in names if n not in bound]
    assert not unbound, f"declared in include/{header} but unknown to the ctypes bindings: {unbound}"


def test_no_cpu_fallback():
    """without a device the compute entry points return an error (PB2_ERR_NO_DEVICE) — nothing
    is computed on the host"""
    import numpy as np
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    L = capi.lib()
    x = np.zeros(8)
    rc = L.pb2_weighted_sum(x.ctypes.data, x.ctypes.data, 1.0, 1.0, x.ctypes.data, 8, None)
    assert rc != 0
    assert b"device" in L.pb2_last_error().lower()
