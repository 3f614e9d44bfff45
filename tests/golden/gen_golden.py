"""Generates the golden fixtures by running the UNMODIFIED reference classes (imported from /root/reference)
on weights / inputs produced by the oracle's deterministic generators.  Run in the build container only:

    python tests/golden/gen_golden.py [resnet] [vis] [kmeans] [metrics] [vit] [spatial]

The fixtures travel with the repo; /root/reference does not exist on the GPU box.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")


def gen_resnet():
    from oracle import resnet50_oracle as O
    from src.resnet import resnet50  # the reference
    torch.manual_seed(0)
    sd = O.make_state_dict(0)
    ref = resnet50(pretrained=False).eval()
    ref.load_state_dict(sd, strict=True)
    patches = O.make_patches(7, 2)
    with torch.no_grad():
        x = O.preprocess(patches)
        feat = ref.forward_extract(x)
        # the same through the reference's own preprocessing objects (torchvision transforms)
        from torchvision import transforms
        tv = torch.nn.Sequential(transforms.ConvertImageDtype(torch.float),
                                 transforms.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225]))
        x_tv = torch.stack([tv(p.permute(2, 0, 1)) for p in patches])
        assert torch.equal(x_tv, x), "preprocessing restatement differs from torchvision transforms"
        feat64 = ref.double().forward_extract(x.double())
    np.savez_compressed(os.path.join(HERE, "resnet50_golden.npz"), weights_seed=0, patches_seed=7, n=2,
                        features=feat.numpy(), features_fp64=feat64.numpy())
    print("resnet golden:", feat.shape, float(feat.abs().mean()))


def gen_vis():
    """Reference ViS (src/tformer_lin.py) on the oracle's seeded weights:
    cfg1 = BASELINE configs[0] (1 slide, 100x2048 -> 1000 genes, depth 6, 16 heads): forward only;
    small = depth 2, 257 genes, batch 3: forward, MSE loss, autograd gradients, 3 AdamW steps (src/vit.py:163-180)."""
    from oracle import vis_oracle as V
    from src.tformer_lin import ViS  # the reference
    out = {}
    # ---- config 1: forward
    sd = V.make_state_dict(0, 1000)
    ref = ViS(num_outputs=1000, input_dim=2048, depth=6, nheads=16, dimensions_f=64, dimensions_s=64, dimensions_c=64,
              device="cpu")
    assert list(ref.state_dict().keys()) == list(sd.keys())
    ref.load_state_dict(sd, strict=True)
    x, _ = V.make_inputs(0, 1, 1000)
    with torch.no_grad():
        out["cfg1_pred"] = ref(x).numpy()
        out["cfg1_pred_fp64"] = ref.double()(x.double()).numpy()
    # ---- small training case (D = 1024 exercises the UNI feature width, 257 genes an unaligned head)
    for tag, D, G, B, depth in (("small", 1024, 257, 3, 2), ("wide", 2048, 1000, 4, 1)):
        sd = V.make_state_dict(1, G, input_dim=D, depth=depth)
        ref = ViS(num_outputs=G, input_dim=D, depth=depth, nheads=16, dimensions_f=64, dimensions_s=64, dimensions_c=64,
                  device="cpu")
        ref.load_state_dict(sd, strict=True)
        opt = torch.optim.AdamW(list(ref.parameters()), lr=1e-3, amsgrad=False, weight_decay=0.)
        loss_fn = torch.nn.MSELoss()
        losses = []
        for step in range(3):
            x, y = V.make_inputs(10 + step, B, G, input_dim=D)
            pred = ref(x)
            loss = loss_fn(pred, y)
            opt.zero_grad()
            loss.backward()
            if step == 0:
                out[f"{tag}_pred0"] = pred.detach().numpy()
                names = [n for n, _ in ref.named_parameters()]
                out[f"{tag}_grad_norms"] = np.array([float(p.grad.double().norm()) for p in ref.parameters()])
                keep = ["pos_emb1D", "transformer.layers.0.0.mixers.3.f.weight", "transformer.layers.0.0.mixers.3.s.weight",
                        "transformer.layers.0.0.mixers.5.c.weight", "transformer.layers.0.0.mixers.5.c.bias",
                        "transformer.layers.0.0.mixers.7.local_norm.weight", "transformer.layers.0.0.mixers.7.summary_norm.bias",
                        "transformer.layers.0.0.projection.bias", "transformer.layers.0.1.net.0.weight",
                        "transformer.layers.0.1.net.1.bias", "linear_head.0.weight", "linear_head.1.bias"]
                for k in keep:
                    out[f"{tag}_grad::{k}"] = dict(ref.named_parameters())[k].grad.numpy().copy()
                assert names == list(sd.keys())
            opt.step()
            losses.append(float(loss))
        out[f"{tag}_losses"] = np.array(losses)
        x, _ = V.make_inputs(99, B, G, input_dim=D)
        with torch.no_grad():
            out[f"{tag}_pred_after3"] = ref(x).numpy()
    np.savez_compressed(os.path.join(HERE, "vis_golden.npz"), **out)
    print("vis golden:", {k: v.shape for k, v in out.items() if "grad::" not in k})


KMEANS_CASES = [  # (tag, slide_id, n, d, modes)
    ("s0", 0, 4096, 2048, 150), ("s1", 1, 4096, 2048, 150), ("s2", 2, 4096, 2048, 150),
    ("uni", 3, 4000, 1024, 120), ("odd", 4, 1237, 512, 110), ("tail", 5, 4099, 256, 130), ("big", 6, 8192, 128, 150),
]


def gen_kmeans():
    """sklearn.cluster.KMeans(n_clusters=100, random_state=0).fit(X).labels_ (the reference's call,
    pre_processing/kmean_features.py:96) on synthetic slides, plus the script's per-label means (:99-105)."""
    import sklearn
    from sklearn.cluster import KMeans
    from oracle import kmeans_oracle as K
    out = {"sklearn_version": np.array(sklearn.__version__)}
    for tag, sid, n, d, modes in KMEANS_CASES:
        X = K.make_slide_features(sid, n=n, d=d, modes=modes)
        km = KMeans(n_clusters=100, random_state=0).fit(X)
        labels = km.labels_
        feats = []
        for pos in range(100):                                   # kmean_features.py:99-105, verbatim semantics
            feats.append(np.mean(X[np.where(labels == pos)], axis=0))
        feats = np.asarray(feats)
        lo, idx, it = K.fit_labels(X)
        assert np.array_equal(lo, labels), f"oracle restatement differs from sklearn on {tag}"
        assert np.array_equal(K.fit_labels(X, pot_mode="emulated")[0], labels), f"emulated BLAS order differs on {tag}"
        assert it == km.n_iter_
        out[f"{tag}_labels"] = labels.astype(np.int16)
        out[f"{tag}_seed_rows"] = idx.astype(np.int32)
        out[f"{tag}_n_iter"] = np.array(km.n_iter_)
        out[f"{tag}_means_checksum"] = np.array([feats.astype(np.float64).sum(), np.abs(feats.astype(np.float64)).sum()])
        out[f"{tag}_means_head"] = feats[:, :8].copy()
        print("kmeans golden", tag, n, d, "iters", km.n_iter_, "sizes", np.bincount(labels).min(), np.bincount(labels).max())
    np.savez_compressed(os.path.join(HERE, "kmeans_golden.npz"), **out)
    gen_kmeans_extra()


# (tag, slide id, n, d, modes, k): other --num_clusters values (kmean_features.py:19) = other local-trial counts 2 + int(ln k)
KMEANS_K_CASES = [("k20", 11, 1500, 256, 40, 20), ("k50", 12, 2000, 128, 80, 50), ("k200", 13, 3000, 256, 260, 200), ("k400", 14, 4096, 128, 500, 400)]
# (tag, slide id, n, d, modes, k, duplicated initial centres): sklearn's `init=X[rows], n_init=1` with duplicate rows ->
# empty clusters in the first iteration(s) -> _relocate_empty_clusters_dense
KMEANS_RELOC_CASES = [("r1", 21, 1500, 64, 150, 100, 1), ("r3", 22, 2000, 128, 150, 100, 3), ("r7", 23, 1237, 256, 150, 100, 7),
                      ("r2", 24, 3000, 64, 150, 50, 2), ("r12", 25, 2500, 128, 150, 200, 12)]


def kmeans_reloc_rows(sid, n, k, ndup):
    rs = np.random.RandomState(sid)
    rows = rs.choice(n, k, replace=False)
    for j in rs.choice(k, ndup, replace=False):
        rows[j] = rows[(j + 1 + rs.randint(k - 1)) % k]
    return rows.astype(np.int32)


def gen_kmeans_extra():
    import warnings
    from sklearn.cluster import KMeans
    from oracle import kmeans_oracle as K
    out = {}
    for tag, sid, n, d, modes, k in KMEANS_K_CASES:
        X = K.make_slide_features(sid, n=n, d=d, modes=modes)
        km = KMeans(n_clusters=k, random_state=0).fit(X)
        lo, idx, it = K.fit_labels(X, k=k)
        assert np.array_equal(lo, km.labels_) and it == km.n_iter_, f"oracle differs from sklearn on {tag}"
        assert np.array_equal(K.fit_labels(X, k=k, pot_mode="emulated")[0], km.labels_), f"emulated BLAS order differs on {tag}"
        out[f"{tag}_labels"] = km.labels_.astype(np.int16); out[f"{tag}_n_iter"] = np.array(km.n_iter_); out[f"{tag}_seed_rows"] = idx.astype(np.int32)
        print("kmeans golden", tag, "k", k, "trials", K.N_LOCAL_TRIALS(k), "iters", km.n_iter_)
    for tag, sid, n, d, modes, k, ndup in KMEANS_RELOC_CASES:
        X = K.make_slide_features(sid, n=n, d=d, modes=modes)
        rows = kmeans_reloc_rows(sid, n, k, ndup)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            km = KMeans(n_clusters=k, init=X[rows], n_init=1).fit(X)
        Xc = np.array(X, dtype=np.float32, copy=True)
        tol = np.float32(np.mean(np.var(Xc, axis=0)) * 1e-4)
        Xc -= Xc.mean(axis=0)
        lo, it = K.lloyd(Xc, Xc[rows].copy(), tol)
        assert np.array_equal(lo, km.labels_) and it == km.n_iter_, f"oracle relocation differs from sklearn on {tag}"
        out[f"{tag}_labels"] = km.labels_.astype(np.int16); out[f"{tag}_n_iter"] = np.array(km.n_iter_); out[f"{tag}_rows"] = rows
        print("kmeans relocation golden", tag, "dups", ndup, "iters", km.n_iter_, "clusters used", len(np.unique(km.labels_)))
    np.savez_compressed(os.path.join(HERE, "kmeans_extra_golden.npz"), **out)


def gen_metrics():
    """The reference's own compute_correlations (src/he2rna.py:140-149), extracted with ast (the module imports tkinter),
    and sklearn's mean_absolute_error (src/vit.py:167) on synthetic batches."""
    import ast
    from sklearn.metrics import mean_absolute_error
    from oracle import metrics_oracle as MO
    src = open("/root/reference/src/he2rna.py").read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "compute_correlations"][0]
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "he2rna.py", "exec"), ns)
    # the reference's own smape (src/vit.py:32-33), extracted the same way (src.vit imports he2rna -> tkinter)
    vsrc = open("/root/reference/src/vit.py").read()
    sfn = [n for n in ast.parse(vsrc).body if isinstance(n, ast.FunctionDef) and n.name == "smape"][0]
    exec(compile(ast.Module(body=[sfn], type_ignores=[]), "vit.py", "exec"), ns)
    out = {}
    for tag, seed, b, g in (("cfg3", 0, 32, 20530), ("small", 1, 5, 301), ("b2", 2, 2, 64)):
        y, p = MO.make_batch(seed, b, g)
        out[f"{tag}_corr"] = np.array(ns["compute_correlations"](y, p))
        out[f"{tag}_mae"] = np.array(mean_absolute_error(y, p))
        out[f"{tag}_smape"] = np.array(ns["smape"](y, p), dtype=np.float64)
        assert abs(MO.smape(y, p) - out[f"{tag}_smape"]) <= 1e-12 * abs(out[f"{tag}_smape"])
        assert abs(MO.compute_correlations(y, p) - out[f"{tag}_corr"]) < 1e-12
        print("metrics golden", tag, float(out[f"{tag}_corr"]), float(out[f"{tag}_mae"]))
    np.savez_compressed(os.path.join(HERE, "metrics_golden.npz"), **out)


def gen_vit():
    """Reference ViT (src/vit.py:93-116) on the oracle's seeded weights.  `src.vit` imports `src.he2rna` (tkinter, wandb,
    h5py: absent) only for `compute_correlations`, which the model classes never touch: that module is stubbed.
    main = the shape src/main.py builds (dim 2048, 16 heads, mlp 2048) at depth 2, 300 genes, batch 2: forward, loss,
    gradients, 3 AdamW steps; small = dim 256, 4 heads, mlp 512, depth 3, 37 tokens, 129 genes, batch 3."""
    import types
    stub = types.ModuleType("src.he2rna")
    stub.compute_correlations = lambda *a, **k: None
    sys.modules["src.he2rna"] = stub
    from oracle import vit_oracle as T
    from src.vit import ViT  # the reference
    out = {}
    for tag, D, H, F_, G, B, depth, N in (("main", 2048, 16, 2048, 300, 2, 2, 100), ("small", 256, 4, 512, 129, 3, 3, 37)):
        sd = T.make_state_dict(2, G, dim=D, depth=depth, heads=H, mlp_dim=F_, num_clusters=N)
        ref = ViT(num_outputs=G, dim=D, depth=depth, heads=H, mlp_dim=F_, dim_head=64, num_clusters=N, device="cpu")
        assert list(ref.state_dict().keys()) == list(sd.keys())
        ref.load_state_dict(sd, strict=True)
        opt = torch.optim.AdamW(list(ref.parameters()), lr=1e-3, amsgrad=False, weight_decay=0.)
        loss_fn = torch.nn.MSELoss()
        losses = []
        for step in range(3):
            x, y = T.make_inputs(20 + step, B, G, input_dim=D, num_clusters=N)
            pred = ref(x)
            loss = loss_fn(pred, y)
            opt.zero_grad()
            loss.backward()
            if step == 0:
                out[f"{tag}_pred0"] = pred.detach().numpy()
                out[f"{tag}_grad_norms"] = np.array([float(p.grad.double().norm()) for p in ref.parameters()])
                for k in ("transformer.layers.0.0.norm.weight", "transformer.layers.0.0.to_qkv.weight", "transformer.layers.1.0.to_out.weight",
                          "transformer.layers.0.1.net.1.bias", "linear_head.1.bias"):
                    g = dict(ref.named_parameters())[k].grad.numpy()
                    out[f"{tag}_grad::{k}"] = g[:64].copy()                 # leading rows keep the fixture small
                with torch.no_grad():
                    ref64 = ViT(num_outputs=G, dim=D, depth=depth, heads=H, mlp_dim=F_, dim_head=64, num_clusters=N, device="cpu").double()
                    ref64.load_state_dict(T.to_double(sd))
                    out[f"{tag}_pred0_fp64"] = ref64(x.double()).numpy()
            opt.step()
            losses.append(float(loss))
        out[f"{tag}_losses"] = np.array(losses)
        x, _ = T.make_inputs(99, B, G, input_dim=D, num_clusters=N)
        with torch.no_grad():
            out[f"{tag}_pred_after3"] = ref(x).numpy()
    np.savez_compressed(os.path.join(HERE, "vit_golden.npz"), **out)
    print("vit golden:", {k: v.shape for k, v in out.items()})


def spatial_case():
    """Synthetic tissue grid shared by the generator and tests/test_spatial_cpu.py: 23 x 19 tile grid with holes."""
    import pandas as pd
    rs = np.random.RandomState(7)
    ps = 256
    pts = [(c, r) for c in range(23) for r in range(19) if rs.rand() < 0.8 and not (8 <= c < 12 and r < 6)]
    df = pd.DataFrame([(c * ps + 1024, r * ps + 512) for c, r in pts], columns=["xcoord", "ycoord"])
    df["xcoord_tf"] = ((df["xcoord"] - min(df["xcoord"])) / ps).astype(int)          # visualize.py:213-214
    df["ycoord_tf"] = ((df["ycoord"] - min(df["ycoord"])) / ps).astype(int)
    return df, ps


def spatial_tile_feature(col, row, dim):
    g = torch.Generator().manual_seed(int(col) * 7919 + int(row))
    return torch.relu(torch.randn(dim, generator=g)) * 0.5


def gen_spatial():
    """The reference's own `sliding_window_method` (spatial_vis/visualize.py:35-102, extracted with ast: the module imports
    openslide and timm) driven by stub `slide` / `transforms_` / feature model and the reference ViS / ViT classes on CPU."""
    import ast
    import types
    from einops import rearrange
    stub = types.ModuleType("src.he2rna")
    stub.compute_correlations = lambda *a, **k: None
    sys.modules["src.he2rna"] = stub
    from oracle import vis_oracle as V
    from oracle import vit_oracle as T
    from src.tformer_lin import ViS
    from src.vit import ViT
    src = open("/root/reference/spatial_vis/visualize.py").read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "sliding_window_method"][0]
    D = 64

    class Patch:
        def __init__(self, col, row):
            self.col, self.row = col, row

        def convert(self, mode):
            return self

    class Slide:
        def read_region(self, loc, level, size):
            return Patch(*loc)

    class Feat:
        def forward_extract(self, x):
            return x

    ns = {"np": np, "torch": torch, "tqdm": lambda it: it, "rearrange": rearrange, "slide": Slide(),
          "transforms_": lambda p: spatial_tile_feature(p.col, p.row, D)}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "visualize.py", "exec"), ns)
    df, ps = spatial_case()
    vis = ViS(num_outputs=9, input_dim=D, depth=1, nheads=2, dimensions_f=64, dimensions_s=64, dimensions_c=64, device="cpu")
    vis.load_state_dict(V.make_state_dict(11, 9, input_dim=D, depth=1, nheads=2))
    vit = ViT(num_outputs=9, dim=D, depth=1, heads=2, mlp_dim=128, dim_head=64, device="cpu")
    vit.load_state_dict(T.make_state_dict(12, 9, dim=D, depth=1, heads=2, mlp_dim=128))
    genes = [0, 3, 8]
    out = {}
    for name, model in (("vis", vis.eval()), ("vit", vit.eval())):
        for stride in (1, 4, 10):
            preds = ns["sliding_window_method"](df=df, patch_size_resized=ps, feat_model=Feat(), model=model, inds_gene_of_interest=genes,
                                                stride=stride, feat_model_type="resnet", feat_dim=D, model_type=name, device="cpu")
            ks = list(preds[genes[0]].keys())
            out[f"{name}_s{stride}_keys"] = np.array(ks, dtype=np.int64)
            out[f"{name}_s{stride}_vals"] = np.array([[preds[g][k] for g in genes] for k in ks], dtype=np.float32)
            print("spatial golden", name, stride, len(ks))
    np.savez_compressed(os.path.join(HERE, "spatial_golden.npz"), **out)


if __name__ == "__main__":
    what = sys.argv[1:] or ["resnet", "vis", "kmeans"]
    for w in what:
        globals()["gen_" + w]()
