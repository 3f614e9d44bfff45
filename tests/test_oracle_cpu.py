"""CPU: the oracle restatements against the golden fixtures produced by the unmodified reference."""
import os
import sys

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def test_resnet_oracle_matches_reference_golden():
    from oracle import resnet50_oracle as O
    g = np.load(os.path.join(GOLD, "resnet50_golden.npz"))
    sd = O.make_state_dict(int(g["weights_seed"]))
    patches = O.make_patches(int(g["patches_seed"]), int(g["n"]))
    with torch.no_grad():
        feat = O.forward_extract(sd, O.preprocess(patches)).numpy()
    assert feat.shape == (2, 2048)
    assert _rel(feat, g["features"]) < 1e-5
    assert _rel(feat, g["features_fp64"]) < 1e-5


def test_resnet_avgpool_is_top_left_7x7():
    # AvgPool2d(7) on the 8x8 map averages rows/cols 0..6 only
    x = torch.arange(2 * 3 * 8 * 8, dtype=torch.float32).view(2, 3, 8, 8)
    a = torch.nn.functional.avg_pool2d(x, 7).reshape(2, 3)
    assert torch.allclose(a, x[:, :, :7, :7].mean((2, 3)))
    assert not torch.allclose(a, x.mean((2, 3)))


def test_vis_oracle_forward_matches_reference_golden_config1():
    """BASELINE configs[0]: 1 slide, 100x2048 -> 1000 genes, CPU."""
    from oracle import vis_oracle as V
    g = np.load(os.path.join(GOLD, "vis_golden.npz"))
    sd = V.make_state_dict(0, 1000)
    assert len(sd) == 1013                                   # 1 013 tensors at depth 6 / 16 heads
    x, _ = V.make_inputs(0, 1, 1000)
    with torch.no_grad():
        pred = V.forward(sd, x).numpy()
    assert pred.shape == (1, 1000)
    assert _rel(pred, g["cfg1_pred"]) < 2e-6
    assert _rel(pred, g["cfg1_pred_fp64"]) < 2e-6


def test_vis_oracle_train_steps_match_reference_golden():
    from oracle import vis_oracle as V
    g = np.load(os.path.join(GOLD, "vis_golden.npz"))
    for tag, D, G, B, depth in (("small", 1024, 257, 3, 2), ("wide", 2048, 1000, 4, 1)):
        sd = V.make_state_dict(1, G, input_dim=D, depth=depth)
        x, y = V.make_inputs(10, B, G, input_dim=D)
        loss, pred, grads = V.loss_and_grads(sd, x, y)
        assert _rel(pred.numpy(), g[f"{tag}_pred0"]) < 2e-6
        norms = np.array([float(grads[k].double().norm()) for k in sd])
        assert np.allclose(norms, g[f"{tag}_grad_norms"], rtol=1e-4, atol=1e-9)
        for key in g.files:
            if key.startswith(f"{tag}_grad::"):
                assert _rel(grads[key.split("::")[1]].numpy(), g[key]) < 1e-5, key
        batches = [V.make_inputs(10 + s, B, G, input_dim=D) for s in range(3)]
        losses = V.train_steps(sd, batches)
        assert np.allclose(losses, g[f"{tag}_losses"], rtol=1e-5)
        x, _ = V.make_inputs(99, B, G, input_dim=D)
        with torch.no_grad():
            after = V.forward(sd, x).numpy()
        assert _rel(after, g[f"{tag}_pred_after3"]) < 1e-4


def test_kmeans_blas_order_restatement_matches_numpy():
    """The fp32 summation orders the CUDA seeding reproduces (oracle/kmeans_oracle.py) against numpy's own BLAS calls."""
    from oracle import kmeans_oracle as K
    rng = np.random.RandomState(0)
    for n in (100, 333, 1000, 4000, 4095, 4096, 4099, 8192, 9999):
        w = np.ones(n, np.float32)
        M = (np.abs(rng.randn(6, n)) * rng.rand(6, n) * 50).astype(np.float32)
        ref = (M @ w.reshape(-1, 1))[:, 0]
        got = np.array([K.blas_order_gemv_row(M[t], t, 6) for t in range(6)], np.float32)
        assert np.array_equal(ref, got), n
        assert (M[:1] @ w)[0] == K.blas_order_sdot(M[0]), n


def test_kmeans_blas_order_other_row_counts():
    """sgemv row grouping for every local-trial count 2 + int(ln k) can take (2 .. 10): groups of four (kind 0), a pair
    (kind 1), a single row (kind 0 order)."""
    from oracle import kmeans_oracle as K
    rng = np.random.RandomState(1)
    for T in range(2, 11):
        for n in (1237, 4096, 4100, 5003):
            M = (rng.rand(T, n) * 3).astype(np.float32)
            ref = (M @ np.ones((n, 1), np.float32))[:, 0]
            got = np.array([K.blas_order_gemv_row(M[t], t, T) for t in range(T)], np.float32)
            assert np.array_equal(ref, got), (T, n)


def test_kmeans_relocation_restatement_matches_sklearn():
    """oracle.relocate_empty_clusters against scikit-learn's own _relocate_empty_clusters_dense, the numpy pairwise-sum
    restatement the CUDA kernel follows, and a full duplicate-initialised fit against the golden."""
    from sklearn.cluster._k_means_common import _relocate_empty_clusters_dense
    from oracle import kmeans_oracle as K
    rs = np.random.RandomState(3)
    for n in (5, 8, 17, 100, 128, 129, 300, 1000, 1024, 2048, 2050):
        a = rs.rand(n).astype(np.float32)
        assert K.pairwise_sum_f32(a) == a.reshape(1, -1).sum(axis=1)[0], n
    for trial in range(10):
        n, d, k = 800, 48, 40
        X = rs.randn(n, d).astype(np.float32); co = rs.randn(k, d).astype(np.float32)
        labels = rs.randint(0, k - 1 - trial % 6, size=n).astype(np.int32)
        w = np.bincount(labels, minlength=k).astype(np.float32)
        cn = np.zeros((k, d), np.float32)
        for i, l in enumerate(labels):
            cn[l] += X[i]
        a_cn, a_w, b_cn, b_w = cn.copy(), w.copy(), cn.copy(), w.copy()
        _relocate_empty_clusters_dense(X, np.ones(n, np.float32), co, a_cn, a_w, labels)
        K.relocate_empty_clusters(X, co, b_cn, b_w, labels)
        assert np.array_equal(a_cn, b_cn) and np.array_equal(a_w, b_w), trial
    sys.path.insert(0, GOLD)
    from gen_golden import kmeans_reloc_rows
    g = np.load(os.path.join(GOLD, "kmeans_extra_golden.npz"))
    tag, sid, n, d, modes, k, ndup = "r3", 22, 2000, 128, 150, 100, 3
    X = K.make_slide_features(sid, n=n, d=d, modes=modes)
    rows = kmeans_reloc_rows(sid, n, k, ndup)
    assert np.array_equal(rows, g[f"{tag}_rows"])
    Xc = np.array(X, dtype=np.float32, copy=True)
    tol = np.float32(np.mean(np.var(Xc, axis=0)) * 1e-4)
    Xc -= Xc.mean(axis=0)
    lab, it = K.lloyd(Xc, Xc[rows].copy(), tol)
    assert np.array_equal(lab, g[f"{tag}_labels"]) and it == int(g[f"{tag}_n_iter"])
    labels, _, it = K.fit_labels(K.make_slide_features(11, n=1500, d=256, modes=40), k=20)
    assert np.array_equal(labels, g["k20_labels"]) and it == int(g["k20_n_iter"])


def test_kmeans_oracle_matches_sklearn_golden_and_live():
    from oracle import kmeans_oracle as K
    g = np.load(os.path.join(GOLD, "kmeans_golden.npz"))
    for tag, sid, n, d, modes in (("odd", 4, 1237, 512, 110), ("tail", 5, 4099, 256, 130)):
        X = K.make_slide_features(sid, n=n, d=d, modes=modes)
        for mode in ("blas", "emulated"):
            labels, idx, it = K.fit_labels(X, pot_mode=mode)
            assert np.array_equal(labels, g[f"{tag}_labels"]) and np.array_equal(idx, g[f"{tag}_seed_rows"]), (tag, mode)
            assert it == int(g[f"{tag}_n_iter"])
        assert np.array_equal(K.cluster_means(X, labels)[:, :8], g[f"{tag}_means_head"])
    from sklearn.cluster import KMeans                       # live check on one more slide (sklearn is the reference's dependency)
    X = K.make_slide_features(21, n=700, d=128, modes=90)
    assert np.array_equal(K.fit_labels(X)[0], KMeans(n_clusters=100, random_state=0).fit(X).labels_)


def test_uni_oracle_structure_matches_torchvision_vit_l_16():
    """UNI parity is unpinned (no timm, no weights); the restatement is cross-checked against torchvision's ViT-L/16,
    which shares the block structure (cls-first tokens, pre-LN, LN eps 1e-6, exact GELU) when LayerScale gamma = 1."""
    from torchvision.models import vit_l_16
    from oracle import uni_oracle as U
    assert len(U.param_shapes()) == 4 + 24 * 14 + 2
    sd = U.make_state_dict(0, depth=3)
    for k in sd:
        if k.endswith("gamma"):
            sd[k] = torch.ones_like(sd[k])
    tv = vit_l_16(weights=None)
    tv.encoder.layers = tv.encoder.layers[:3]
    tv.heads = torch.nn.Identity()
    missing = tv.load_state_dict(U.to_torchvision(sd), strict=True)
    x = U.preprocess(U.make_patches(0, 2))
    with torch.no_grad():
        want = tv.eval()(x).numpy()
        got = U.forward(sd, x).numpy()
    assert got.shape == (2, 1024)
    assert _rel(got, want) < 1e-5


def test_resize_oracle_matches_pillow():
    """oracle/resize_oracle.py (Resample.c restated) against Pillow through the reference's own transform object
    (pre_processing/compute_features_hdf5.py:53-56: transforms.Resize(224) on Image.fromarray(tile).convert("RGB"))."""
    pytest = __import__("pytest")
    Image = pytest.importorskip("PIL.Image")
    from torchvision import transforms
    from oracle import resize_oracle as R
    rs = np.random.RandomState(0)
    tf = transforms.Resize(224)
    for hw in ((256, 256), (300, 300), (224, 224), (512, 512), (256, 320), (250, 250), (200, 200)):
        for kind in range(2):
            a = (rs.rand(*hw, 3) * 255).astype(np.uint8) if kind == 0 else np.clip(np.cumsum(rs.randn(*hw, 3), axis=1) * 8 + 128, 0, 255).astype(np.uint8)
            ref = np.asarray(tf(Image.fromarray(a).convert("RGB")))
            assert np.array_equal(R.resize(a, 224), ref), hw
