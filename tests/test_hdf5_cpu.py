"""CPU: the built-in HDF5 codec against a file written by the HDF5 library itself, a strict structural
checker of the on-disk invariants libhdf5 relies on (applied to the library's file AND to ours), and round trips of the
reference's two layouts (patch_gen_hdf5.py:119-120, compute_features_hdf5.py:134-135, kmean_features.py:75,108)."""
import os
import struct

import numpy as np
import pytest

from sequoia_pub_b200 import hdf5

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "hdf5_libhdf5_double.mat")
UNDEF = 0xFFFFFFFFFFFFFFFF


def check_structure(path):
    """Independent walk of a version-0-superblock file; returns {name: (shape, dtype message bytes, data address, size)}."""
    d = open(path, "rb").read()
    base = 0
    while d[base:base + 8] != b"\x89HDF\r\n\x1a\n":
        base = 512 if base == 0 else base * 2
        assert base < len(d)
    sb = d[base:]
    assert sb[8] == 0 and sb[13] == 8 and sb[14] == 8
    leaf_k, int_k = struct.unpack_from("<HH", sb, 16)
    base_addr, _, eof, _ = struct.unpack_from("<4Q", sb, 24)
    assert base_addr == base
    name_off, root, cache, _, bt, heap = struct.unpack_from("<QQIIQQ", sb, 56)
    assert name_off == 0 and cache == 1

    def at(a, n):
        assert base + a + n <= len(d), "address beyond the end of the file"
        return d[base + a:base + a + n]

    def messages(a):
        ver, nmsg, ref, hsize = struct.unpack_from("<BxHII", at(a, 12))
        assert ver == 1 and ref == 1 and hsize % 8 == 0
        blk, p, out = at(a + 16, hsize), 0, []
        for _ in range(nmsg):
            t, s, fl = struct.unpack_from("<HHB", blk, p)
            assert s % 8 == 0 and p + 8 + s <= hsize
            assert t != 0x10, "continuation blocks are not expected in these files"
            out.append((t, fl, blk[p + 8:p + 8 + s]))
            p += 8 + s
        assert p == hsize, "messages must tile the header exactly"
        return out

    sym = [b for t, _, b in messages(root) if t == 0x11]
    assert len(sym) == 1 and struct.unpack_from("<QQ", sym[0]) == (bt, heap), "root entry cache == symbol table message"
    assert at(heap, 4) == b"HEAP" and at(heap, 8)[4] == 0
    hsize, hfree, hdata = struct.unpack_from("<QQQ", at(heap, 32), 8)
    seg = at(hdata, hsize)
    assert hsize % 8 == 0 and seg[:1] == b"\0"
    while hfree != 1:                                   # free list: (next, size) blocks inside the segment
        nxt, fsz = struct.unpack_from("<QQ", seg, hfree)
        assert fsz >= 16 and hfree + fsz <= hsize
        hfree = nxt

    def name(off):
        return seg[off:seg.index(b"\0", off)]

    entries, levels = [], {}

    def walk(a, lo, hi, expect_level=None):
        hdr = at(a, 24)
        assert hdr[:4] == b"TREE" and hdr[4] == 0
        level, used = hdr[5], struct.unpack_from("<H", hdr, 6)[0]
        left, right = struct.unpack_from("<QQ", hdr, 8)
        assert expect_level is None or level == expect_level
        assert used <= 2 * int_k
        at(a, 24 + (4 * int_k + 1) * 8)                 # the full node must be allocated
        levels.setdefault(level, []).append((a, left, right))
        body = at(a + 24, 8 + 16 * used)
        keys = [struct.unpack_from("<Q", body, 16 * i)[0] for i in range(used + 1)]
        if used:
            assert name(keys[0]) == lo and name(keys[-1]) == hi
        for i in range(used):
            child = struct.unpack_from("<Q", body, 8 + 16 * i)[0]
            klo, khi = name(keys[i]), name(keys[i + 1])
            assert klo < khi
            if level:
                walk(child, klo, khi, level - 1)
                continue
            s = at(child, 8 + 2 * leaf_k * 40)
            assert s[:4] == b"SNOD" and s[4] == 1
            n = struct.unpack_from("<H", s, 6)[0]
            assert 1 <= n <= 2 * leaf_k
            names = []
            for j in range(n):
                no, obj, ctype = struct.unpack_from("<QQI", s, 8 + 40 * j)
                names.append(name(no))
                entries.append((name(no).decode(), obj))
                assert ctype == 0
            assert names == sorted(names) and klo < names[0] and names[-1] == khi

    hdr = at(bt, 24)
    used = struct.unpack_from("<H", hdr, 6)[0]
    if used:
        body = at(bt + 24, 8 + 16 * used)
        walk(bt, name(struct.unpack_from("<Q", body, 0)[0]), name(struct.unpack_from("<Q", body, 16 * used)[0]))
    for nodes in levels.values():                       # sibling chains, left to right
        for i, (a, left, right) in enumerate(nodes):
            assert left == (nodes[i - 1][0] if i else UNDEF)
            assert right == (nodes[i + 1][0] if i + 1 < len(nodes) else UNDEF)
    assert [e[0].encode() for e in entries] == sorted(e[0].encode() for e in entries)

    out = {}
    for nm, obj in entries:
        msgs = {t: b for t, _, b in messages(obj)}
        assert {1, 3, 8} <= set(msgs)
        sp = msgs[1]
        assert sp[0] == 1
        shape = struct.unpack_from(f"<{sp[1]}Q", sp, 8)
        lay = msgs[8]
        if lay[0] == 3:
            assert lay[1] == 1
            addr, size = struct.unpack_from("<QQ", lay, 2)
        else:
            assert lay[0] in (1, 2) and lay[2] == 1
            addr = struct.unpack_from("<Q", lay, 8)[0]
            size = int(np.prod(shape)) * struct.unpack_from("<I", msgs[3], 4)[0]
        if size:
            at(addr, size)
        out[nm] = (shape, bytes(msgs[3]), addr, size)
    return out, eof, base, len(d)


def test_reader_decodes_a_file_written_by_libhdf5():
    entries, _, base, _ = check_structure(GOLDEN)
    assert base == 512 and list(entries) == ["testdouble"]
    with hdf5.File(GOLDEN) as f:
        assert f.keys() == ["testdouble"] and "testdouble" in f and "nope" not in f
        ds = f["testdouble"]
        assert ds.shape == (9, 1) and ds.dtype == np.dtype("<f8")
        assert np.array_equal(ds[:].ravel(), np.linspace(0, 2 * np.pi, 9))
        with pytest.raises(KeyError):
            f["nope"]


def test_writer_emits_the_same_datatype_message_as_libhdf5():
    entries, *_ = check_structure(GOLDEN)
    lib = entries["testdouble"][1]
    assert hdf5._encode_dtype(np.float64) == lib[:len(hdf5._encode_dtype(np.float64))]
    for dt in ("<f4", "<f8", "<f2", "u1", "<i4", "<u2", "<i8", ">f4", ">i2"):
        assert hdf5._decode_dtype(hdf5._encode_dtype(dt)) == np.dtype(dt)


@pytest.mark.parametrize("n_tiles", [0, 1, 8, 9, 256, 257, 2100])
def test_patch_file_round_trip_and_key_order(tmp_path, n_tiles):
    rs = np.random.RandomState(n_tiles)
    coords = [(int(x), int(y)) for x, y in zip(rs.permutation(n_tiles) * 7 % 5000, rs.randint(0, 90000, n_tiles))]
    tiles = {f"{x}_{y}": rs.randint(0, 256, (8, 8, 3)).astype(np.uint8) for x, y in coords}
    p = tmp_path / "slide.hdf5"
    with hdf5.File(p, "w") as f:
        for k, v in tiles.items():                      # insertion order is not name order
            f.create_dataset(k, data=v)
    entries, eof, _, size = check_structure(p)
    assert eof == size and len(entries) == len(tiles)
    with hdf5.File(p, "r") as f:
        keys = list(f.keys())
        assert keys == sorted(tiles, key=str.encode) and len(f) == len(tiles)
        for k in keys[:50]:
            assert f[k].shape == (8, 8, 3) and f[k].dtype == np.uint8 and np.array_equal(f[k][:], tiles[k])
        out = np.empty((len(keys), 8, 8, 3), np.uint8)
        f.read_many(keys, out)
        assert all(np.array_equal(out[i], tiles[k]) for i, k in enumerate(keys))
        sub = keys[::3][::-1]                           # random.sample order (compute_features_hdf5.py:112-113)
        f.read_many(sub, out)
        assert all(np.array_equal(out[i], tiles[k]) for i, k in enumerate(sub))
        with pytest.raises(ValueError):
            f.create_dataset("x", data=np.zeros(3))


def test_feature_file_append_in_place(tmp_path):
    rs = np.random.RandomState(0)
    feats = rs.rand(321, 2048).astype(np.float32)
    p = tmp_path / "WSI.h5"
    f = hdf5.File(p, "w")                               # compute_features_hdf5.py:134-136
    f.create_dataset("resnet_features", data=feats)
    f.close()
    with pytest.raises(ValueError):
        f["resnet_features"]
    check_structure(p)
    f = hdf5.File(p, "r+")                              # kmean_features.py:75-108
    h = f["resnet_features"]
    assert h.shape[0] == 321 and "cluster_features" not in f.keys()
    assert np.array_equal(np.asarray(h), feats) and np.array_equal(h[np.where(feats[:, 0] > 0.5)], feats[feats[:, 0] > 0.5])
    means = rs.rand(100, 2048).astype(np.float32)
    f.create_dataset("cluster_features", data=means)
    with pytest.raises(ValueError):
        f.create_dataset("cluster_features", data=means)
    f.close()
    entries, eof, _, size = check_structure(p)
    assert eof == size and sorted(entries) == ["cluster_features", "resnet_features"]
    with hdf5.File(p, "r") as f:                        # src/read_data.py:47-49, src/utils.py:30-33
        assert list(f.keys()) == ["cluster_features", "resnet_features"]
        assert np.array_equal(f["cluster_features"][:], means) and np.array_equal(f["resnet_features"][:], feats)
        assert f["cluster_features"][:].dtype == np.float32


def test_append_to_the_library_file_and_error_paths(tmp_path):
    p = tmp_path / "lib.h5"
    p.write_bytes(open(GOLDEN, "rb").read())
    root_before = struct.unpack_from("<Q", open(GOLDEN, "rb").read(), 512 + 64)[0]
    with hdf5.File(p, "r+") as f:
        f.create_dataset("cluster_features", data=np.arange(12, dtype=np.float32).reshape(3, 4))
    entries, eof, base, size = check_structure(p)                # still a consistent file, user block and base address kept
    assert base == 512 and base + eof == size and sorted(entries) == ["cluster_features", "testdouble"]
    assert struct.unpack_from("<Q", p.read_bytes(), 512 + 64)[0] == root_before      # the library's root header is re-linked in place
    with hdf5.File(p, "r") as f:
        assert f.keys() == ["cluster_features", "testdouble"]
        assert np.array_equal(f["testdouble"][:].ravel(), np.linspace(0, 2 * np.pi, 9))
        assert np.array_equal(f["cluster_features"][:], np.arange(12, dtype=np.float32).reshape(3, 4))
    bad = tmp_path / "bad.h5"
    bad.write_bytes(b"not an hdf5 file" * 10)
    with pytest.raises(OSError):
        hdf5.File(bad)
    with pytest.raises(OSError):
        hdf5.File(tmp_path / "missing.h5")
    trunc = tmp_path / "trunc.h5"
    trunc.write_bytes(open(GOLDEN, "rb").read()[:3000])
    with pytest.raises(OSError):
        hdf5.File(trunc)["testdouble"][:]
    empty = tmp_path / "empty.h5"
    with hdf5.File(empty, "w") as f:
        f.create_dataset("e", data=np.zeros((0, 2048), np.float32))
    with hdf5.File(empty) as f:
        assert f["e"].shape == (0, 2048) and f["e"][:].shape == (0, 2048)


def test_open_file_falls_back_to_the_builtin_codec(tmp_path):
    f = hdf5.open_file(tmp_path / "a.h5", "w")
    try:
        import h5py  # noqa: F401
        assert not isinstance(f, hdf5.File)
    except ImportError:
        assert isinstance(f, hdf5.File)
    f.close()
