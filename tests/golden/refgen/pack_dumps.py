"""Pack the raw U.<cycle>.bin dumps written by burgers_dump_main.cpp into one .npz.

Test infrastructure only.  Layout of the dumps is documented in burgers_dump_main.cpp.
The .npz holds, per dumped cycle c:  U_c [nblocks, ncomp, nk, nj, ni] float64,
plus block metadata (gid, level, lx1, lx2, lx3 — tree-relative), block bounds
(xmin[3], xmax[3]), times and dts.
"""
import glob
import os
import sys

import numpy as np


def read_dump(path):
    with open(path, "rb") as f:
        hdr = np.frombuffer(f.read(28), dtype="<i4")
        assert hdr[0] == 0x50423230, "bad magic"
        nb, nc, nk, nj, ni, cycle = (int(x) for x in hdr[1:])
        time, dt = np.frombuffer(f.read(16), dtype="<f8")
        meta = np.zeros((nb, 5), dtype=np.int32)
        bounds = np.zeros((nb, 6), dtype=np.float64)
        data = np.zeros((nb, nc, nk, nj, ni), dtype=np.float64)
        n = nc * nk * nj * ni
        for b in range(nb):
            meta[b] = np.frombuffer(f.read(20), dtype="<i4")
            bounds[b] = np.frombuffer(f.read(48), dtype="<f8")
            data[b] = np.frombuffer(f.read(8 * n), dtype="<f8").reshape(nc, nk, nj, ni)
    return cycle, time, dt, meta, bounds, data


def main(src_dir, out):
    files = sorted(glob.glob(os.path.join(src_dir, "U.*.bin")),
                   key=lambda p: int(p.split(".")[-2]))
    arrays = {}
    times, dts, cycles = [], [], []
    for p in files:
        cycle, time, dt, meta, bounds, data = read_dump(p)
        arrays[f"U_{cycle}"] = data
        arrays["meta"] = meta
        arrays["bounds"] = bounds
        if os.environ.get("PB2_PER_CYCLE_META"):  # adaptive meshes: the block list changes
            arrays[f"meta_{cycle}"] = meta
            arrays[f"bounds_{cycle}"] = bounds
        cycles.append(cycle)
        times.append(time)
        dts.append(dt)
    arrays["cycles"] = np.array(cycles, dtype=np.int32)
    arrays["times"] = np.array(times)
    arrays["dts"] = np.array(dts)
    np.savez_compressed(out, **arrays)
    print(out, {k: v.shape for k, v in arrays.items()})


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
