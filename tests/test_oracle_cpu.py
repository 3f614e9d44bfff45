"""CPU: the oracle restatements against the golden fixtures produced by the unmodified reference."""
import os

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
