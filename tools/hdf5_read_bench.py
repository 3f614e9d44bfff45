"""Host-side read throughput of a patch file through the built-in HDF5 codec: 4096 uint8 [256,256,3] tiles
(805 MB) written with `hdf5.File`, then read back (page cache warm) with one preadv per tile into one buffer (`read_many`)
and, for comparison, tile by tile through dataset objects + np.stack as the reference loop does (compute_features_hdf5.py:117)."""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequoia_pub_b200 import hdf5  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rs = np.random.RandomState(0)
tile = rs.randint(0, 256, (256, 256, 3)).astype(np.uint8)
with tempfile.TemporaryDirectory() as d:
    path = os.path.join(d, "slide.hdf5")
    t0 = time.perf_counter()
    with hdf5.File(path, "w") as f:
        for i in range(n):
            f.create_dataset(f"{(i % 64) * 256}_{(i // 64) * 256}", data=np.roll(tile, i, axis=0))
    t_write = time.perf_counter() - t0
    size = os.path.getsize(path)
    out = np.empty((n, 256, 256, 3), np.uint8)
    t_bulk = {}
    for thr in (1, 2, 4, 8, 1, 8):
        t0 = time.perf_counter()
        with hdf5.File(path, "r") as f:
            keys = list(f.keys())
            t_open = time.perf_counter() - t0
            f.read_many(keys, out, threads=thr)
        t_bulk[thr] = time.perf_counter() - t0
    t0 = time.perf_counter()
    with hdf5.File(path, "r") as f:
        stacked = np.stack([f[k][:] for k in f.keys()])
    t_loop = time.perf_counter() - t0
    assert np.array_equal(out, stacked)
    print(f"{n} tiles, file {size / 1e6:.0f} MB: write {t_write:.2f} s ({size / t_write / 1e9:.2f} GB/s); open+index {t_open * 1e3:.0f} ms; "
          f"read_many " + ", ".join(f"{k} thr {v:.3f} s ({size / v / 1e9:.2f} GB/s)" for k, v in t_bulk.items()) + f"; per-dataset loop + np.stack {t_loop:.3f} s ({size / t_loop / 1e9:.2f} GB/s)")
