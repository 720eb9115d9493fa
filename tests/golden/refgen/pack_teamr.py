"""Pack the dumps of teamr_dump_main.cpp (file index = 3 * cycle + field) into a compact fixture:
per cycle c the block list (meta_c, bounds_c) and one CRC-32 per block and field (crc_c_f,
f = 0 face, 1 edge, 2 node), plus the array shapes.  Test infrastructure only."""
import glob
import os
import sys
import zlib

import numpy as np

from pack_dumps import read_dump


def main(src_dir, out):
    files = glob.glob(os.path.join(src_dir, "U.*.bin"))
    ncyc = len(files) // 3
    arrays = {"ncycles": np.array(ncyc - 1)}
    for c in range(ncyc):
        for f in range(3):
            _, _, _, meta, bounds, data = read_dump(os.path.join(src_dir, f"U.{3 * c + f}.bin"))
            arrays[f"meta_{c}"] = meta
            arrays[f"bounds_{c}"] = bounds
            arrays[f"shape_{c}_{f}"] = np.array(data.shape)
            arrays[f"crc_{c}_{f}"] = np.array(
                [zlib.crc32(np.ascontiguousarray(data[b]).tobytes()) for b in range(data.shape[0])],
                dtype=np.uint32)
    np.savez_compressed(out, **arrays)
    print(out, ncyc, "dumps,", [int(arrays[f"meta_{c}"].shape[0]) for c in range(ncyc)], "blocks")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
