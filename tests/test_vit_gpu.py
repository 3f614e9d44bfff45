"""ViT softmax-attention baseline: CUDA path (through the C ABI) vs the oracle restatement and the golden
vectors produced by the reference's own `src/vit.py::ViT`.  Parity bars as for ViS: predictions and every
per-parameter gradient L2-relative and max-abs/max-scale <= 1e-4 vs the fp32 reference; 3-step AdamW trajectory <= 1e-3."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-4
CASES = {"main": dict(D=2048, H=16, F=2048, G=300, B=2, depth=2, N=100), "small": dict(D=256, H=4, F=512, G=129, B=3, depth=3, N=37)}


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _maxrel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def _sd(c, seed=2):
    from oracle import vit_oracle as T
    return T.make_state_dict(seed, c["G"], dim=c["D"], depth=c["depth"], heads=c["H"], mlp_dim=c["F"], num_clusters=c["N"])


def _model(sd, c, device="cuda"):
    from sequoia_pub_b200.vit import ViT
    m = ViT(num_outputs=c["G"], dim=c["D"], depth=c["depth"], heads=c["H"], mlp_dim=c["F"], dim_head=64, num_clusters=c["N"], device=device)
    m.load_state_dict(sd, strict=True)
    return m.to(device)


def _inputs(c, seed):
    from oracle import vit_oracle as T
    return T.make_inputs(seed, c["B"], c["G"], input_dim=c["D"], num_clusters=c["N"])


# ------------------------------------------------------------------------------------------------ CPU
def test_vit_oracle_matches_reference_golden():
    """The restatement (oracle/vit_oracle.py) against outputs of the reference class: forward, gradients, 3 AdamW steps."""
    from oracle import vit_oracle as T
    g = np.load(os.path.join(GOLD, "vit_golden.npz"))
    for tag, c in CASES.items():
        sd = _sd(c)
        x, y = _inputs(c, 20)
        loss, pred, grads = T.loss_and_grads(sd, x, y)
        assert _rel(pred, g[f"{tag}_pred0"]) < 2e-6 and _rel(pred, g[f"{tag}_pred0_fp64"]) < 1e-5
        norms = np.array([float(grads[k].double().norm()) for k in sd])
        assert np.allclose(norms, g[f"{tag}_grad_norms"], rtol=1e-4, atol=1e-9)
        for key in g.files:
            if key.startswith(f"{tag}_grad::"):
                assert _rel(grads[key.split("::")[1]][:64], g[key]) < 1e-5, key
        losses = T.train_steps(sd, [_inputs(c, 20 + s) for s in range(3)])
        assert np.allclose(losses, g[f"{tag}_losses"], rtol=1e-5)
        with torch.no_grad():
            after = T.forward(sd, _inputs(c, 99)[0])
        assert _rel(after, g[f"{tag}_pred_after3"]) < 1e-4


def test_state_dict_schema_and_layout():
    """Reference names/shapes (src/vit.py:93-105), strict load, and the C layout: aligned, contiguous per stage, no CPU path."""
    import ctypes as C
    from oracle import vit_oracle as T
    from sequoia_pub_b200 import _lib
    from sequoia_pub_b200.vit import ViT
    m = ViT(num_outputs=1000, dim=2048, depth=6, heads=16, mlp_dim=2048, dim_head=64)
    sd = T.make_state_dict(0, 1000)
    assert list(m.state_dict().keys()) == T.param_names(6) and len(sd) == 1 + 10 * 6 + 4
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    m.load_state_dict(sd, strict=True)
    cfg = m._config()
    table, total = m._layout_table(cfg)
    assert table == sorted(table) and table[0] == 0 and all(o % 64 == 0 for o in table)
    n_params = sum(p.numel() for p in m.parameters())
    assert n_params <= total < n_params + 64 * len(table)
    m._cfg = cfg
    rng = m._stage_ranges()
    assert rng[0][0] == 0 and rng[-1][1] == total and all(rng[i][1] == rng[i + 1][0] for i in range(len(rng) - 1)) and len(rng) == 7
    L = _lib.lib()
    assert L.sq_vit_act_bytes(C.byref(cfg), 32) > 0 and L.sq_vit_bwd_bytes(C.byref(cfg), 32) > 0
    for bad, msg in ((_lib.VitConfig(2000, 6, 16, 100, 10, 2048), b"multiple of 64"), (_lib.VitConfig(2048, 6, 16, 129, 10, 2048), b"num_clusters"),
                     (_lib.VitConfig(2048, 6, 16, 100, 10, 100), b"mlp_dim")):
        assert L.sq_vit_param_table_len(C.byref(bad)) < 0 and msg in L.sq_last_error()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 100, 2048))
    with pytest.raises(NotImplementedError):
        m.transformer(torch.zeros(1, 100, 2048))
    with pytest.raises(NotImplementedError):
        ViT(num_outputs=10, dim=128, depth=1, heads=2, mlp_dim=128, dim_head=32)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("tag", list(CASES))
def test_forward_and_gradients_match_oracle_and_golden(tag):
    from oracle import vit_oracle as T
    c = CASES[tag]
    g = np.load(os.path.join(GOLD, "vit_golden.npz"))
    sd = _sd(c)
    m = _model(sd, c).train()
    x, y = _inputs(c, 20)
    xg = x.cuda().requires_grad_(True)
    pred = m(xg)
    loss = torch.nn.MSELoss()(pred, y.cuda())
    loss.backward()
    assert _rel(pred, g[f"{tag}_pred0"]) < TOL and _maxrel(pred, g[f"{tag}_pred0"]) < TOL
    assert abs(loss.item() - g[f"{tag}_losses"][0]) / g[f"{tag}_losses"][0] < 1e-5
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xc = x.clone().requires_grad_(True)
    lo = torch.nn.functional.mse_loss(T.forward(params, xc), y)
    grads = torch.autograd.grad(lo, list(params.values()) + [xc])
    ref = dict(zip(list(params.keys()) + ["x"], grads))
    worst = ("", 0.0)
    for name, p in m.named_parameters():
        assert p.grad is not None, name
        e = max(_rel(p.grad, ref[name]), _maxrel(p.grad, ref[name]))
        if e > worst[1]:
            worst = (name, e)
    print(f"\n[vit parity] {tag}: pred {_rel(pred, g[f'{tag}_pred0']):.3e}; worst per-parameter gradient error {worst[1]:.3e} ({worst[0]}); "
          f"dx {_rel(xg.grad, ref['x']):.3e}")
    assert worst[1] < TOL, worst
    assert _rel(xg.grad, ref["x"]) < TOL
    norms = np.array([p.grad.double().norm().item() for p in m.parameters()])
    assert np.allclose(norms, g[f"{tag}_grad_norms"], rtol=2e-4, atol=1e-9)
    for key in g.files:
        if key.startswith(f"{tag}_grad::"):
            assert _rel(dict(m.named_parameters())[key.split("::")[1]].grad[:64], g[key]) < TOL, key
    with torch.no_grad():                                       # inference path (no autograd graph) gives the same numbers
        assert torch.equal(m.eval()(x.cuda()), pred.detach())


@pytest.mark.gpu
@pytest.mark.parametrize("opt_kind", ["torch", "fused", "trainer"])
@pytest.mark.parametrize("tag", list(CASES))
def test_adamw_trajectory_matches_golden(opt_kind, tag):
    """3 steps of src/vit.py:163-180 with AdamW(lr=1e-3, wd=0): torch.optim.AdamW on the drop-in module, FusedAdamW, FusedTrainer."""
    from sequoia_pub_b200.tformer_lin import FusedAdamW
    from sequoia_pub_b200.train import FusedTrainer
    c = CASES[tag]
    g = np.load(os.path.join(GOLD, "vit_golden.npz"))
    m = _model(_sd(c), c).train()
    losses = []
    if opt_kind == "trainer":
        tr = FusedTrainer(m, lr=1e-3, weight_decay=0.0)
        for s in range(3):
            x, y = _inputs(c, 20 + s)
            losses.append(tr.step(x.cuda(), y.cuda()).item())
    else:
        opt = (torch.optim.AdamW(list(m.parameters()), lr=1e-3, amsgrad=False, weight_decay=0.) if opt_kind == "torch"
               else FusedAdamW(list(m.parameters()), lr=1e-3, amsgrad=False, weight_decay=0.))
        for s in range(3):
            x, y = _inputs(c, 20 + s)
            loss = torch.nn.MSELoss()(m(x.cuda()), y.cuda())
            opt.zero_grad()
            loss.backward()
            opt.step()
            losses.append(loss.item())
    assert np.allclose(losses, g[f"{tag}_losses"], rtol=2e-4), (losses, g[f"{tag}_losses"])
    m.eval()
    with torch.no_grad():
        after = m(_inputs(c, 99)[0].cuda())
    e = _rel(after, g[f"{tag}_pred_after3"])
    print(f"\n[vit parity] {tag}/{opt_kind}: losses {losses}, pred after 3 steps L2-rel {e:.3e}")
    assert e < 1e-3


@pytest.mark.gpu
def test_full_shape_train_step_edge_batches_and_determinism():
    """The shape src/main.py trains (dim 2048, depth 6, 16 heads, mlp 2048, 100 tokens) at batch 32 -> 1000 genes against the
    oracle's forward on a 2-slide subset; batch 1 and an empty batch; bit-identical repeats; head replacement (main.py:155-157)."""
    import time
    from oracle import vit_oracle as T
    from sequoia_pub_b200.train import FusedTrainer
    c = dict(D=2048, H=16, F=2048, G=1000, B=32, depth=6, N=100)
    sd = _sd(c, seed=3)
    m = _model(sd, c).eval()
    x, y = _inputs(c, 30)
    with torch.no_grad():
        p1 = m(x.cuda()); p2 = m(x.cuda())
        want = T.forward(sd, x[:2])
        assert torch.equal(p1, p2)
        assert _rel(p1[:2], want) < TOL and _maxrel(p1[:2], want) < TOL
        assert torch.equal(m(x[:1].cuda()), p1[:1])
        assert m(x[:0].cuda()).shape == (0, 1000)
    tr = FusedTrainer(m.train(), lr=1e-3)
    l0 = tr.step(x.cuda(), y.cuda()).item()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        tr.step(x.cuda(), y.cuda())
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"\n[vit timing] train step batch 32, depth 6, 1000 genes: {dt * 1e3:.2f} ms -> {32 / dt:.0f} slides/s")
    want0 = torch.nn.functional.mse_loss(T.forward(sd, x[:4]), y[:4]).item()          # loss on a 4-slide subset for scale only
    assert l0 > 0 and abs(np.log(l0 / want0)) < 0.5
    m.linear_head = torch.nn.Sequential(torch.nn.LayerNorm(2048), torch.nn.Linear(2048, 77)).cuda()
    with torch.no_grad():
        assert m.eval()(x[:2].cuda()).shape == (2, 77)
    with pytest.raises(ValueError):
        m(torch.zeros(1, 99, 2048, device="cuda"))
