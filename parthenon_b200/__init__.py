"""parthenon_b200 — B200-native (sm_100a) implementation of Parthenon's per-cycle ghost-zone
hot path: boundary pack/unpack, prolongation/restriction ghost fill and the
benchmarks/burgers stencil, behind a C ABI (include/parthenon_b200.h).

`capi` binds libpb200.so (kernels + C ABI); `host` binds libpb200_host.so (the C++ host
framework mirroring Parthenon's package / MeshData / task-list API).
"""
__version__ = "0.1.0"
