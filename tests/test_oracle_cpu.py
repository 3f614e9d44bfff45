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
