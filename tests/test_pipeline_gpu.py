"""End-to-end slice of BASELINE configs[4] on one GPU: synthetic tiles -> ResNet-50 features -> k-means(100) cluster features
-> one ViS train step, each stage checked against its oracle on the same inputs (stage-wise, as SURVEY §8d prescribes)."""
import os

import numpy as np
import pytest
import torch


@pytest.mark.gpu
def test_extract_reduce_train_slice():
    from oracle import kmeans_oracle as K
    from oracle import resnet50_oracle as RO
    from oracle import vis_oracle as V
    from sequoia_pub_b200 import pipeline
    from sequoia_pub_b200.resnet import resnet50
    from sequoia_pub_b200.tformer_lin import ViS
    from sequoia_pub_b200.train import FusedTrainer
    sd = RO.make_state_dict(0)
    model = resnet50().eval()
    model.load_state_dict(sd)
    model = model.cuda()
    slides = []
    for sid in range(2):
        rs = np.random.RandomState(sid)
        modes = rs.randint(0, 256, size=(12, 1, 1, 3))
        tiles = np.clip(modes[rs.randint(0, 12, size=160)] + rs.randint(-40, 40, size=(160, 256, 256, 3)), 0, 255).astype(np.uint8)
        feats = pipeline.extract_tiles(model, tiles, batch_size=64)
        assert feats.shape == (160, 2048) and feats.dtype == np.float32
        with torch.no_grad():
            want = RO.forward_extract(sd, RO.preprocess(torch.from_numpy(tiles[:4]))).numpy()
        assert np.linalg.norm(feats[:4] - want) / np.linalg.norm(want) < 5e-3
        cf = pipeline.reduce_features(feats, 100)
        labels, _, _ = K.fit_labels(feats)                       # same features -> the oracle's labels and means
        assert np.array_equal(cf, K.cluster_means(feats, labels))
        slides.append(cf)
        assert pipeline.reduce_features(feats[:50], 100) is None  # fewer tiles than clusters: skipped like the reference
    x = torch.from_numpy(np.stack(slides))                        # [2, 100, 2048]
    G = 64
    vsd = V.make_state_dict(0, G, depth=1)
    m = ViS(num_outputs=G, input_dim=2048, depth=1, nheads=16, dimensions_f=64, dimensions_s=64, dimensions_c=64)
    m.load_state_dict(vsd)
    m = m.cuda().train()
    y = torch.rand(2, G) * 10
    loss = FusedTrainer(m, lr=1e-3).step(x.cuda(), y.cuda()).item()
    want = V.train_steps(vsd, [(x, y)])[0]
    assert abs(loss - want) / want < 1e-4


def test_run_slides_isolates_failures_and_shards(capsys):
    from sequoia_pub_b200.pipeline import run_slides
    seen = []

    def fn(s):
        if s == "bad":
            raise RuntimeError("unreadable slide")
        seen.append(s)
    ids = ["a", "bad", "b", "c", "d"]
    assert run_slides(ids, fn, rank=0, world=2) == ["a", "b"]
    assert run_slides(ids, fn, rank=1, world=2) == ["c", "d"]
    assert "unreadable slide" in capsys.readouterr().out


@pytest.mark.gpu
def test_file_level_drivers_with_the_builtin_hdf5_codec(tmp_path):
    """compute_features_hdf5.py:110-136 then kmean_features.py:75-108 on real files, no h5py involved."""
    import random
    from oracle import resnet50_oracle as RO
    from sequoia_pub_b200 import hdf5, pipeline
    from sequoia_pub_b200.resnet import resnet50
    model = resnet50().eval()
    model.load_state_dict(RO.make_state_dict(0))
    model = model.cuda()
    rs = np.random.RandomState(5)
    modes = rs.randint(0, 256, size=(12, 1, 1, 3))
    tiles = {f"{256 * (i % 13)}_{256 * (i // 13)}": np.clip(modes[rs.randint(12)] + rs.randint(-40, 40, size=(256, 256, 3)), 0, 255).astype(np.uint8)
             for i in range(150)}
    patch_file = tmp_path / "patches" / "S1" / "S1.hdf5"
    os.makedirs(patch_file.parent)
    with hdf5.File(patch_file, "w") as f:                       # patch_gen_hdf5.py:119-120
        for k, v in tiles.items():
            f.create_dataset(k, data=v)
    feature_file = tmp_path / "features" / "TCGA-X" / "S1" / "S1.h5"
    feats = pipeline.extract_slide(model, patch_file, feature_file, prefer_h5py=False)
    keys = sorted(tiles, key=str.encode)
    want = pipeline.extract_tiles(model, np.stack([tiles[k] for k in keys]))
    assert np.array_equal(feats, want)                          # rows in key order
    sub = pipeline.extract_slide(model, patch_file, tmp_path / "sub.h5", max_patch_number=100, rng=random.Random(3), prefer_h5py=False)
    pick = random.Random(3).sample(keys, 100)
    assert np.array_equal(sub, want[[keys.index(k) for k in pick]])
    cf = pipeline.reduce_slide(feature_file, prefer_h5py=False)
    assert cf.shape == (100, 2048) and np.array_equal(cf, pipeline.reduce_features(want, 100))
    assert pipeline.reduce_slide(feature_file, prefer_h5py=False) is None          # already there: skipped (:91-94)
    with hdf5.File(feature_file) as f:
        assert list(f.keys()) == ["cluster_features", "resnet_features"]
        assert np.array_equal(f["cluster_features"][:], cf) and np.array_equal(f["resnet_features"][:], want)
