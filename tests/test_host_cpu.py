"""CPU: host-side logic of the extraction driver that mirrors pre_processing/compute_features_hdf5.py:110-113."""
import random


def test_select_keys_matches_the_reference_subsampling():
    from sequoia_pub_b200.extract import select_keys
    keys = [f"{x}_{y}" for x in range(0, 9000, 256) for y in range(0, 9000, 256)]      # tile datasets are named "{x}_{y}"
    assert select_keys(keys, max_patch_number=len(keys)) == keys                        # all tiles, key order kept
    random.seed(99)                                                                     # the script seeds `random` (:41)
    want = random.sample(keys, 400)                                                     # :112-113
    random.seed(99)
    assert select_keys(keys, max_patch_number=400) == want
    rng = random.Random(5)
    assert select_keys(keys, 10, rng) == random.Random(5).sample(keys, 10)
    assert select_keys([], 10) == []


def test_product_modules_refuse_to_run_without_a_gpu():
    """No CPU fallback anywhere: every product entry point raises instead of computing on the host."""
    import numpy as np
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from sequoia_pub_b200 import metrics
    from sequoia_pub_b200.kmeans import KMeans
    from sequoia_pub_b200.resnet import resnet50
    with pytest.raises(RuntimeError):
        resnet50().eval().forward_extract(torch.zeros(1, 3, 256, 256))
    with pytest.raises((RuntimeError, AssertionError)):
        KMeans(n_clusters=100, random_state=0).fit(np.zeros((200, 64), np.float32))
    with pytest.raises(RuntimeError):
        metrics.step_metrics(torch.zeros(4, 8), torch.zeros(4, 8))


def test_device_slide_dataset_semantics():
    """f-2: preloaded data path keeps the reference's sample order, drops unreadable samples like custom_collate_fn, and
    shards global batches evenly across ranks (runs on CPU tensors: it is indexing only)."""
    import torch
    from sequoia_pub_b200.data import DeviceSlideDataset
    g = torch.Generator().manual_seed(0)
    samples = [(torch.randn(100, 16, generator=g) if i != 3 else None, torch.randn(7, generator=g), f"wsi{i}", "TCGA-X") for i in range(10)]
    ds = DeviceSlideDataset(samples, device="cpu")
    assert len(ds) == 9 and ds.dropped == ["wsi3"] and ds.num_genes == 7 and ds.feature_dim == 16
    got = list(ds.batches(4))
    assert [b[0].shape[0] for b in got] == [4, 4, 1]
    assert got[0][2] == ["wsi0", "wsi1", "wsi2", "wsi4"]
    assert torch.equal(got[0][0][3], samples[4][0]) and torch.equal(got[0][1][3], samples[4][1])
    e1 = [b[2] for b in ds.batches(4, shuffle=True, generator=torch.Generator().manual_seed(1))]
    e2 = [b[2] for b in ds.batches(4, shuffle=True, generator=torch.Generator().manual_seed(1))]
    assert e1 == e2 and sorted(sum(e1, [])) == sorted(ds.names)
    r0 = list(ds.batches(4, rank=0, world=2)); r1 = list(ds.batches(4, rank=1, world=2))
    assert [b[0].shape[0] for b in r0] == [2, 2] and r0[0][2] + r1[0][2] == got[0][2]


def test_device_dataset_from_feature_files(tmp_path, capsys):
    """f-1 + f-2: `read_samples` restates SuperTileRNADataset.__getitem__ (src/read_data.py:38-56) over feature files written
    by the built-in codec: rna_* column gather, '.svs' stripped from TCGA paths, unreadable slides -> None -> dropped."""
    import numpy as np
    import pandas as pd
    import torch
    from sequoia_pub_b200 import hdf5
    from sequoia_pub_b200.data import DeviceSlideDataset, read_samples
    rs = np.random.RandomState(0)
    names = ["TCGA-A.svs", "TCGA-B", "TCGA-C", "GTEX-D"]
    df = pd.DataFrame({"wsi_file_name": names, "tcga_project": ["P1", "P1", "P2", "P2"], "patient_id": list("abcd"),
                       "rna_G1": rs.rand(4), "rna_G2": rs.rand(4), "other": rs.rand(4), "rna_G3": rs.rand(4)})
    feats = {}
    for name, proj in zip(names, df.tcga_project):
        if name == "TCGA-C":
            continue                                              # no feature file: unreadable sample
        stem = name.replace(".svs", "")
        d = tmp_path / proj / stem
        d.mkdir(parents=True)
        feats[name] = rs.rand(100, 32).astype(np.float32)
        with hdf5.File(d / (stem + ".h5"), "w") as f:
            f.create_dataset("resnet_features", data=rs.rand(130, 32).astype(np.float32))
            f.create_dataset("cluster_features", data=feats[name])
    samples = list(read_samples(df, str(tmp_path), prefer_h5py=False))
    assert [s[2] for s in samples] == names and samples[2][0] is None
    assert "TCGA-C" in capsys.readouterr().out
    for (f, r, name, proj), row in zip(samples, df.itertuples()):
        assert np.allclose(r.numpy(), [row.rna_G1, row.rna_G2, row.rna_G3]) and r.dtype == torch.float32
        if f is not None:
            assert np.array_equal(f.numpy(), feats[name])
    ds = DeviceSlideDataset.from_files(df, str(tmp_path), device="cpu", prefer_h5py=False)
    assert len(ds) == 3 and ds.dropped == ["TCGA-C"] and ds.feature_dim == 32 and ds.num_genes == 3
    # src/utils.py:20-40: slides without a feature directory, without a readable file or without the dataset are filtered out
    from sequoia_pub_b200.data import filter_no_features
    (tmp_path / "P2" / "TCGA-E").mkdir()
    (tmp_path / "P2" / "TCGA-E" / "TCGA-E.h5").write_bytes(b"garbage")
    df2 = pd.concat([df, pd.DataFrame({"wsi_file_name": ["TCGA-E"], "tcga_project": ["P2"]})], ignore_index=True)
    df2["wsi_file_name"] = df2["wsi_file_name"].str.replace(".svs", "")
    kept = filter_no_features(df2, str(tmp_path), "cluster_features", prefer_h5py=False)
    assert kept["wsi_file_name"].tolist() == ["TCGA-A", "TCGA-B", "GTEX-D"] and kept.index.tolist() == [0, 1, 2]
    assert filter_no_features(df2, str(tmp_path), "uni_features", prefer_h5py=False).empty


def test_read_tiles_builtin_codec_and_h5py_like_objects(tmp_path):
    """pipeline.read_tiles = `f_read[key][:]` for every key (compute_features_hdf5.py:117): in-place bulk read with the built-in
    codec, per-dataset `read_direct` with anything h5py-shaped; key order (including a random.sample order) is preserved."""
    import numpy as np
    import torch
    from sequoia_pub_b200 import hdf5, pipeline
    rs = np.random.RandomState(1)
    tiles = {f"{x}_{y}": rs.randint(0, 256, (16, 16, 3)).astype(np.uint8) for x in range(0, 1280, 256) for y in range(0, 768, 256)}
    p = tmp_path / "s.hdf5"
    with hdf5.File(p, "w") as f:
        for k, v in tiles.items():
            f.create_dataset(k, data=v)
    with hdf5.File(p) as f:
        keys = list(f.keys())[::-1]
        got = pipeline.read_tiles(f, keys, pinned=False)
        assert isinstance(got, torch.Tensor) and got.dtype == torch.uint8 and tuple(got.shape) == (len(keys), 16, 16, 3)
        assert all(np.array_equal(got[i].numpy(), tiles[k]) for i, k in enumerate(keys))
        assert pipeline.read_tiles(f, []).shape == (0, 256, 256, 3)

    class FakeDataset:
        def __init__(self, a):
            self.a, self.shape = a, a.shape

        def read_direct(self, out):
            out[...] = self.a

    class FakeFile(dict):
        pass
    ff = FakeFile({k: FakeDataset(v) for k, v in tiles.items()})
    got2 = pipeline.read_tiles(ff, keys)
    assert isinstance(got2, np.ndarray) and np.array_equal(got2, got.numpy())


def test_resize_coefficient_tables_match_the_pillow_restatement():
    """Host half of the GPU resize: the library's precompute_coeffs restatement (no GPU needed) vs oracle/resize_oracle.py."""
    import numpy as np
    from oracle import resize_oracle as R
    from sequoia_pub_b200 import preproc
    for a, b in ((256, 224), (300, 224), (512, 224), (200, 224), (224, 224), (257, 224), (1000, 224)):
        bo, ko = R.coeffs(a, b)
        bl, kl = preproc.coeff_tables(a, b)
        assert np.array_equal(bo, bl) and np.array_equal(ko, kl), (a, b)
    assert preproc.resize_size(256, 320) == (224, 280) and preproc.resize_size(512, 256) == (448, 224)


def test_extract_slide_refuses_a_mismatched_feat_type():
    import pytest
    from sequoia_pub_b200 import pipeline
    from sequoia_pub_b200.resnet import resnet50
    with pytest.raises(ValueError):
        pipeline.extract_slide(resnet50(), "nope.hdf5", "out.h5", feat_type="uni")
    with pytest.raises(ValueError):
        pipeline.extract_slide(resnet50(), "nope.hdf5", "out.h5", feat_type="dino")
