"""GPU: the single-token broadcast of `ViS.forward` / `ViT.forward` (what spatial_vis/visualize.py:80 relies on) and the
sliding-window spatial inference on the CUDA aggregators against the dictionaries the reference function produced."""
import os

import numpy as np
import pytest
import torch

from test_spatial_cpu import D, GENES, GG, _features

HERE = os.path.dirname(os.path.abspath(__file__))


def _gpu_models():
    from oracle import vis_oracle as V
    from oracle import vit_oracle as T
    from sequoia_pub_b200.tformer_lin import ViS
    from sequoia_pub_b200.vit import ViT
    vis_sd = V.make_state_dict(11, 9, input_dim=D, depth=1, nheads=2)
    vit_sd = T.make_state_dict(12, 9, dim=D, depth=1, heads=2, mlp_dim=128)
    vis = ViS(num_outputs=9, input_dim=D, depth=1, nheads=2, dimensions_f=64, dimensions_s=64, dimensions_c=64)
    vis.load_state_dict(vis_sd)
    vit = ViT(num_outputs=9, dim=D, depth=1, heads=2, mlp_dim=128, dim_head=64)
    vit.load_state_dict(vit_sd)
    return {"vis": (vis.cuda().eval(), lambda x: V.forward(vis_sd, x)), "vit": (vit.cuda().eval(), lambda x: T.forward(vit_sd, x))}


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["vis", "vit"])
def test_single_token_inputs_broadcast_like_the_reference(name):
    m, oracle = _gpu_models()[name]
    g = torch.Generator().manual_seed(5)
    x2 = torch.relu(torch.randn(100, D, generator=g)) * 0.5                 # the unbatched [100, D] tensor of visualize.py:80
    with torch.no_grad():
        got = m(x2.cuda()).cpu()
        want = oracle(x2)
        assert got.shape == (100, 9)
        assert ((got - want).norm() / want.norm()).item() < 1e-4
        assert torch.equal(m(x2[:7, None, :].cuda()).cpu(), got[:7])        # [B, 1, D] is the same thing
    xg = x2[:3, None, :].clone().cuda().requires_grad_(True)
    m.train()
    m(xg).square().mean().backward()
    xc = x2[:3, None, :].clone().requires_grad_(True)
    oracle(xc).square().mean().backward()
    assert xg.grad.shape == (3, 1, D)
    assert ((xg.grad.cpu() - xc.grad).norm() / xc.grad.norm()).item() < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["vis", "vit"])
@pytest.mark.parametrize("stride", [1, 10])
def test_sliding_window_on_the_cuda_aggregators(name, stride):
    from sequoia_pub_b200.spatial import sliding_window_method
    gold = np.load(os.path.join(HERE, "golden", "spatial_golden.npz"))
    df, ps = GG.spatial_case()
    m, _ = _gpu_models()[name]
    preds = sliding_window_method(df, ps, None, m, GENES, stride, "resnet", D, model_type=name, device="cuda", tile_features=_features(df))
    keys = list(preds[GENES[0]].keys())
    assert keys == list(gold[f"{name}_s{stride}_keys"])
    vals = np.array([[preds[q][k] for q in GENES] for k in keys], dtype=np.float32)
    want = gold[f"{name}_s{stride}_vals"]
    err = np.abs(vals - want).max() / np.abs(want).max()
    print(f"\n[spatial parity] {name} stride {stride}: {len(keys)} tiles, max-rel {err:.2e}")
    assert err < 1e-4
