"""CPU: sliding-window spatial inference (sequoia_pub_b200/spatial.py) against the reference's own `sliding_window_method`
(spatial_vis/visualize.py:35-102).  The golden file holds the dictionaries the reference function returned for the ViS and
ViT classes of the reference on a synthetic tissue grid (tests/golden/gen_golden.py::gen_spatial); here the orchestration
runs with the oracle restatements as the aggregator (no CUDA needed: spatial.py only schedules)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("gen_golden", os.path.join(HERE, "golden", "gen_golden.py"))
GG = importlib.util.module_from_spec(spec)
spec.loader.exec_module(GG)
D, GENES = 64, [0, 3, 8]


def _features(df):
    return torch.stack([GG.spatial_tile_feature(c, r, D) for c, r in zip(df["xcoord"], df["ycoord"])])


def _models():
    from oracle import vis_oracle as V
    from oracle import vit_oracle as T
    vis_sd = V.make_state_dict(11, 9, input_dim=D, depth=1, nheads=2)
    vit_sd = T.make_state_dict(12, 9, dim=D, depth=1, heads=2, mlp_dim=128)
    return {"vis": lambda x: V.forward(vis_sd, x), "vit": lambda x: T.forward(vit_sd, x)}


@pytest.mark.parametrize("name", ["vis", "vit"])
@pytest.mark.parametrize("stride", [1, 4, 10])
def test_matches_the_reference_function(name, stride):
    from sequoia_pub_b200.spatial import sliding_window_method
    g = np.load(os.path.join(HERE, "golden", "spatial_golden.npz"))
    df, ps = GG.spatial_case()
    preds = sliding_window_method(df, ps, None, _models()[name], GENES, stride, "resnet", D, model_type=name, device="cpu",
                                  tile_features=_features(df), windows_per_batch=37)
    keys = list(preds[GENES[0]].keys())
    assert keys == list(g[f"{name}_s{stride}_keys"])                       # same tiles, same insertion order
    vals = np.array([[preds[q][k] for q in GENES] for k in keys], dtype=np.float32)
    want = g[f"{name}_s{stride}_vals"]
    assert vals.shape == want.shape and np.abs(vals - want).max() <= 1e-5 * np.abs(want).max()
    assert all(isinstance(preds[q][k], np.float32) for q in GENES for k in keys[:5])


def _literal(df, out_of_window, stride, genes):
    """visualize.py:43-99 with the aggregator replaced by a table lookup (window origin -> prediction vector)."""
    max_x, max_y = max(df["xcoord_tf"]), max(df["ycoord_tf"])
    preds = {g: {} for g in genes}
    for x in range(0, max_x, stride):
        for y in range(0, max_y, stride):
            window = df[((df["xcoord_tf"] >= x) & (df["xcoord_tf"] < (x + 10))) & ((df["ycoord_tf"] >= y) & (df["ycoord_tf"] < (y + 10)))]
            if window.shape[0] > ((10 * 10) / 2):
                predictions = out_of_window(window.index[0])
                for g in genes:
                    for key in window.index:
                        if stride == 10:
                            preds[g][key] = predictions[g]
                        elif key not in preds[g]:
                            preds[g][key] = [predictions[g]]
                        else:
                            preds[g][key].append(predictions[g])
    if stride < 10:
        for g in genes:
            for key in preds[g]:
                preds[g][key] = np.mean(preds[g][key])
    return preds


@pytest.mark.parametrize("stride", [1, 3, 10, 12])
def test_window_attribution_and_means_are_bit_identical(stride):
    """With identical per-window outputs the batched attribution must equal the reference's dictionary loops bit for bit
    (np.mean over python lists uses pairwise summation from 8 elements on; the vectorised path must keep that order)."""
    from sequoia_pub_b200.spatial import sliding_window_method
    df, ps = GG.spatial_case()
    table = np.random.RandomState(3).randn(len(df), 5).astype(np.float32) * 3 + 1
    model = lambda x: x[:, 0, :5]                                          # "prediction" = first 5 features of the window's first tile
    feats = np.zeros((len(df), D), np.float32)
    feats[:, :5] = table
    got = sliding_window_method(df, ps, None, model, [1, 4], stride, "resnet", D, model_type="vis", device="cpu", tile_features=feats)
    want = _literal(df, lambda first: table[first], stride, [1, 4])
    for g in (1, 4):
        assert list(got[g].keys()) == list(want[g].keys())
        for k in want[g]:
            if stride > 10:
                assert got[g][k] == want[g][k]
            else:
                assert got[g][k] == want[g][k] and type(got[g][k]) is type(want[g][k]), (g, k)


def test_intended_semantics_featurize_and_errors():
    from sequoia_pub_b200.spatial import sliding_window_method, window_index
    df, ps = GG.spatial_case()
    feats = _features(df)
    seen = []

    def featurize(idx):
        seen.extend(idx)
        return feats[idx]
    model = lambda x: x.sum(dim=1)[:, :4]                                  # sees the zero-padded [nw, 100, D] windows
    got = sliding_window_method(df, ps, None, model, [2], 10, "resnet", D, device="cpu", featurize=featurize, reference_semantics=False)
    wins = window_index(df["xcoord_tf"], df["ycoord_tf"], 10)
    assert sorted(seen) == sorted(set(np.concatenate(wins).tolist()))       # every needed tile featurised exactly once
    w0 = wins[0]
    assert np.isclose(got[2][df.index[w0[0]]], feats[w0].sum(0)[2].item(), rtol=1e-6)
    assert all(51 <= len(w) <= 100 for w in wins)
    with pytest.raises(ValueError):
        sliding_window_method(df, ps, None, model, [2], 10, "resnet", D, device="cpu")
    with pytest.raises(NotImplementedError):
        sliding_window_method(df, ps, None, model, [2], 10, "resnet", D, model_type="he2rna", device="cpu", tile_features=feats)
    assert sliding_window_method(df.iloc[:0], ps, None, model, [2], 10, "resnet", D, device="cpu", tile_features=feats[:0]) == {2: {}}
