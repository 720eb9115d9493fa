"""Pack raw U.<cycle>.bin dumps (layout: burgers_dump_main.cpp) of a LARGE run into a compact
fixture: per cycle the block list and, per block, the CRC-32 of its bytes (full extents, ghosts
included) and the sum of its interior cells.  Bit-exact parity can then be checked at sizes
whose raw fields would not fit in the repository.  Test infrastructure only."""
import glob
import os
import sys
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from pack_dumps import read_dump  # noqa: E402


def block_crcs(data):
    return np.array([zlib.crc32(np.ascontiguousarray(data[b]).tobytes()) for b in range(data.shape[0])],
                    dtype=np.uint32)


def main(src_dir, out, ng):
    files = sorted(glob.glob(os.path.join(src_dir, "U.*.bin")), key=lambda p: int(p.split(".")[-2]))
    arrays, times, dts, cycles = {}, [], [], []
    for p in files:
        cycle, time, dt, meta, bounds, data = read_dump(p)
        arrays[f"bounds_{cycle}"] = bounds
        arrays[f"crc_{cycle}"] = block_crcs(data)
        sl = tuple(slice(ng, -ng) if data.shape[2 + d] > 1 else slice(None) for d in range(3))
        arrays[f"sum_{cycle}"] = data[(slice(None), slice(None)) + sl].sum(axis=(1, 2, 3, 4))
        cycles.append(cycle)
        times.append(time)
        dts.append(dt)
    arrays["cycles"] = np.array(cycles, dtype=np.int32)
    arrays["times"] = np.array(times)
    arrays["dts"] = np.array(dts)
    np.savez_compressed(out, **arrays)
    print(out, len(files), "cycles")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]))
