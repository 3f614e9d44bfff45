"""ViS aggregator: CUDA path (through the C ABI) vs the oracle restatement and the reference's golden vectors.

Parity bars: predictions L2-relative and max-abs/max-scale <= 1e-4 vs the fp32 reference; gradients the
same metric per parameter tensor; AdamW trajectory <= 1e-3 after 3 steps."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-4


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _maxrel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def _model(sd, G, D=2048, depth=6, device="cuda"):
    from sequoia_pub_b200.tformer_lin import ViS
    m = ViS(num_outputs=G, input_dim=D, depth=depth, nheads=16, dimensions_f=64, dimensions_s=64, dimensions_c=64, device=device)
    m.load_state_dict(sd, strict=True)
    return m.to(device)


def test_state_dict_schema_matches_reference():
    """1 013 tensors with the reference's names and shapes, strict load both ways (CPU, no kernels)."""
    from oracle import vis_oracle as V
    from sequoia_pub_b200.tformer_lin import ViS
    m = ViS(num_outputs=1000, input_dim=2048, depth=6, nheads=16, dimensions_f=64, dimensions_s=64, dimensions_c=64)
    sd = V.make_state_dict(0, 1000)
    assert list(m.state_dict().keys()) == V.param_names(6, 16) and len(sd) == 1013
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    m.load_state_dict(sd, strict=True)
    assert sum(p.numel() for p in m.parameters()) == 91_229_160
    with pytest.raises(NotImplementedError):
        m.transformer(torch.zeros(1, 100, 2048))
    with pytest.raises(RuntimeError):          # no CPU fallback
        m(torch.zeros(1, 100, 2048))


@pytest.mark.gpu
def test_forward_config1_matches_golden():
    from oracle import vis_oracle as V
    g = np.load(os.path.join(GOLD, "vis_golden.npz"))
    m = _model(V.make_state_dict(0, 1000), 1000).eval()
    x, _ = V.make_inputs(0, 1, 1000)
    with torch.no_grad():
        pred = m(x.cuda())
    e1, e2 = _rel(pred, g["cfg1_pred"]), _maxrel(pred, g["cfg1_pred"])
    print(f"\n[vis parity] config 1 forward: L2-rel {e1:.3e}, max-rel {e2:.3e} (vs fp64 {_rel(pred, g['cfg1_pred_fp64']):.3e})")
    assert pred.shape == (1, 1000) and e1 < TOL and e2 < TOL
    # the 'b ... d -> b (...) d' rearrange: a [B,10,10,D] input is the same 100 tokens
    with torch.no_grad():
        assert torch.equal(m(x.cuda().view(1, 10, 10, 2048)), pred)
        assert torch.equal(m(x.cuda()), pred)           # deterministic


@pytest.mark.gpu
@pytest.mark.parametrize("tag,D,G,B,depth", [("small", 1024, 257, 3, 2), ("wide", 2048, 1000, 4, 1)])
def test_gradients_match_oracle_and_golden(tag, D, G, B, depth):
    from oracle import vis_oracle as V
    g = np.load(os.path.join(GOLD, "vis_golden.npz"))
    sd = V.make_state_dict(1, G, input_dim=D, depth=depth)
    m = _model(sd, G, D, depth).train()
    x, y = V.make_inputs(10, B, G, input_dim=D)
    xg = x.cuda().requires_grad_(True)
    pred = m(xg)
    loss = torch.nn.MSELoss()(pred, y.cuda())
    loss.backward()
    assert _rel(pred, g[f"{tag}_pred0"]) < TOL and _maxrel(pred, g[f"{tag}_pred0"]) < TOL
    assert abs(loss.item() - g[f"{tag}_losses"][0]) / g[f"{tag}_losses"][0] < 1e-5
    # oracle gradients (CPU autograd of the restatement), incl. dL/dx
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xc = x.clone().requires_grad_(True)
    lo = torch.nn.functional.mse_loss(V.forward(params, xc), y)
    grads = torch.autograd.grad(lo, list(params.values()) + [xc])
    ref = dict(zip(list(params.keys()) + ["x"], grads))
    worst = ("", 0.0)
    for name, p in m.named_parameters():
        assert p.grad is not None, name
        e = max(_rel(p.grad, ref[name]), _maxrel(p.grad, ref[name]))
        if e > worst[1]:
            worst = (name, e)
    print(f"\n[vis parity] {tag} gradients: worst per-parameter error {worst[1]:.3e} ({worst[0]}); dx {_rel(xg.grad, ref['x']):.3e}")
    assert worst[1] < TOL, worst
    assert _rel(xg.grad, ref["x"]) < TOL
    norms = np.array([p.grad.double().norm().item() for p in m.parameters()])
    assert np.allclose(norms, g[f"{tag}_grad_norms"], rtol=2e-4, atol=1e-9)
    for key in g.files:
        if key.startswith(f"{tag}_grad::"):
            assert _rel(dict(m.named_parameters())[key.split("::")[1]].grad, g[key]) < TOL, key


@pytest.mark.gpu
@pytest.mark.parametrize("opt_kind", ["torch", "fused", "trainer"])
@pytest.mark.parametrize("tag,D,G,B,depth", [("small", 1024, 257, 3, 2), ("wide", 2048, 1000, 4, 1)])
def test_adamw_trajectory_matches_golden(opt_kind, tag, D, G, B, depth):
    """3 steps of src/vit.py:163-180 with AdamW(lr=1e-3, wd=0): unmodified torch.optim.AdamW on the drop-in module,
    the FusedAdamW optimizer, and the fully fused FusedTrainer."""
    from oracle import vis_oracle as V
    from sequoia_pub_b200.tformer_lin import FusedAdamW
    from sequoia_pub_b200.train import FusedTrainer
    g = np.load(os.path.join(GOLD, "vis_golden.npz"))
    m = _model(V.make_state_dict(1, G, input_dim=D, depth=depth), G, D, depth).train()
    losses = []
    if opt_kind == "trainer":
        tr = FusedTrainer(m, lr=1e-3, weight_decay=0.0)
        for s in range(3):
            x, y = V.make_inputs(10 + s, B, G, input_dim=D)
            losses.append(tr.step(x.cuda(), y.cuda()).item())
    else:
        opt = (torch.optim.AdamW(list(m.parameters()), lr=1e-3, amsgrad=False, weight_decay=0.) if opt_kind == "torch"
               else FusedAdamW(list(m.parameters()), lr=1e-3, amsgrad=False, weight_decay=0.))
        loss_fn = torch.nn.MSELoss()
        for s in range(3):
            x, y = V.make_inputs(10 + s, B, G, input_dim=D)
            pred = m(x.cuda())
            loss = loss_fn(pred, y.cuda())
            opt.zero_grad()
            loss.backward()
            opt.step()
            losses.append(loss.item())
    assert np.allclose(losses, g[f"{tag}_losses"], rtol=2e-4), (losses, g[f"{tag}_losses"])
    x, _ = V.make_inputs(99, B, G, input_dim=D)
    m.eval()
    with torch.no_grad():
        after = m(x.cuda())
    e = _rel(after, g[f"{tag}_pred_after3"])
    print(f"\n[vis parity] {tag}/{opt_kind}: losses {losses}, pred after 3 steps L2-rel {e:.3e}")
    assert e < 1e-3


@pytest.mark.gpu
def test_gradient_accumulation_and_head_replacement():
    """Two backward passes without zero_grad accumulate (autograd semantics); replacing linear_head re-lays-out the
    flat buffer (src/main.py:155-157)."""
    from oracle import vis_oracle as V
    D, G, B, depth = 1024, 257, 2, 1
    sd = V.make_state_dict(3, G, input_dim=D, depth=depth)
    m = _model(sd, G, D, depth).train()
    x, y = V.make_inputs(5, B, G, input_dim=D)
    torch.nn.functional.mse_loss(m(x.cuda()), y.cuda()).backward()
    g1 = [p.grad.clone() for p in m.parameters()]
    torch.nn.functional.mse_loss(m(x.cuda()), y.cuda()).backward()
    for a, p in zip(g1, m.parameters()):
        assert torch.allclose(p.grad, 2 * a, rtol=1e-5, atol=1e-12)
    m.zero_grad()
    torch.manual_seed(0)
    m.linear_head = torch.nn.Sequential(torch.nn.LayerNorm(D), torch.nn.Linear(D, 77)).cuda()
    sd2 = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    with torch.no_grad():
        pred = m(x.cuda())
        want = V.forward(sd2, x)
    assert pred.shape == (B, 77) and _rel(pred, want) < TOL


@pytest.mark.gpu
def test_config3_shape_forward_backward():
    """BASELINE configs[2] shapes: 32 slides, 100x2048 -> 20530 genes, depth 6 (oracle on the CPU takes ~20 s)."""
    from oracle import vis_oracle as V
    G, B = 20530, 32
    sd = V.make_state_dict(2, G)
    m = _model(sd, G).train()
    x, y = V.make_inputs(3, B, G)
    pred = m(x.cuda())
    loss = torch.nn.functional.mse_loss(pred, y.cuda())
    loss.backward()
    lo, pr, grads = V.loss_and_grads(sd, x, y)
    e1, e2 = _rel(pred, pr), _maxrel(pred, pr)
    worst = max(((max(_rel(p.grad, grads[n]), _maxrel(p.grad, grads[n])), n) for n, p in m.named_parameters()))
    print(f"\n[vis parity] config 3: pred L2-rel {e1:.3e} max-rel {e2:.3e}; loss {loss.item():.6f} vs {lo.item():.6f}; "
          f"worst gradient error {worst[0]:.3e} ({worst[1]})")
    assert e1 < TOL and e2 < TOL and worst[0] < TOL
    assert abs(loss.item() - lo.item()) / lo.item() < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("B", [1, 5])
def test_ragged_batches(B):
    from oracle import vis_oracle as V
    D, G, depth = 1024, 130, 1
    sd = V.make_state_dict(4, G, input_dim=D, depth=depth)
    m = _model(sd, G, D, depth).eval()
    x, _ = V.make_inputs(6, B, G, input_dim=D)
    with torch.no_grad():
        assert _rel(m(x.cuda()), V.forward(sd, x)) < TOL


@pytest.mark.gpu
def test_host_batch_feeder_delivers_batches_in_order():
    from sequoia_pub_b200.train import HostBatchFeeder
    g = torch.Generator().manual_seed(0)
    batches = [(torch.randn(3, 100, 64, generator=g).pin_memory(), torch.randn(3, 7, generator=g).pin_memory()) for _ in range(5)]
    seen = []
    for xd, yd in HostBatchFeeder(batches, "cuda"):
        seen.append((xd.cpu().clone(), yd.cpu().clone()))        # consumed on the current stream before the slot is reused
    assert len(seen) == 5
    for (xa, ya), (xb, yb) in zip(seen, batches):
        assert torch.equal(xa, xb) and torch.equal(ya, yb)


@pytest.mark.gpu
def test_error_paths_and_other_shapes():
    """Wrong inputs fail loudly with the library's message; a non-default token count / head count still matches the oracle."""
    from oracle import vis_oracle as V
    from sequoia_pub_b200.tformer_lin import ViS, FusedAdamW
    D, G, depth, H, N = 512, 33, 1, 4, 37
    sd = V.make_state_dict(7, G, input_dim=D, depth=depth, nheads=H, num_clusters=N)
    m = ViS(num_outputs=G, input_dim=D, depth=depth, nheads=H, dimensions_f=64, dimensions_s=64, dimensions_c=64, num_clusters=N)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().train()
    x, y = V.make_inputs(8, 3, G, input_dim=D, num_clusters=N)
    pred = m(x.cuda())
    torch.nn.functional.mse_loss(pred, y.cuda()).backward()
    _, want, grads = V.loss_and_grads(sd, x, y)
    assert _rel(pred, want) < TOL
    assert max(_rel(p.grad, grads[n]) for n, p in m.named_parameters()) < TOL
    with pytest.raises(ValueError):
        m(torch.zeros(2, N + 1, D, device="cuda"))                       # wrong token count
    with pytest.raises(ValueError):
        m(torch.zeros(2, N, D, device="cuda", dtype=torch.float64))      # wrong dtype
    with pytest.raises(NotImplementedError):
        ViS(num_outputs=4, input_dim=D, depth=1, nheads=2, dimensions_f=32, dimensions_s=64, dimensions_c=64)
    with pytest.raises(NotImplementedError):
        FusedAdamW(list(m.parameters()), amsgrad=True)
    opt = FusedAdamW(list(m.parameters()), lr=1e-3, weight_decay=0.0)
    opt.step()
    assert float(next(iter(opt.state.values()))["step"]) == 1.0
    empty = m.eval()(torch.zeros(0, N, D, device="cuda"))
    assert tuple(empty.shape) == (0, G)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["vis", "vit"])
def test_two_forwards_one_backward_accumulate(kind):
    """`loss = f(model(x1)) + f(model(x2))`: both backward calls run before any gradient is accumulated, so each needs its own
    flat gradient buffer (the views of the first are still held by autograd).  Gradients must equal g1 + g2 from separate passes."""
    from oracle import vis_oracle as V
    D, G, B, depth = 1024, 130, 2, 1
    if kind == "vis":
        m = _model(V.make_state_dict(4, G, input_dim=D, depth=depth), G, D, depth).train()
    else:
        from sequoia_pub_b200.vit import ViT
        torch.manual_seed(4)
        m = ViT(num_outputs=G, dim=D, depth=depth, heads=16, mlp_dim=512, dim_head=64).cuda().train()
    x1, y1 = V.make_inputs(6, B, G, input_dim=D)
    x2, y2 = V.make_inputs(7, B, G, input_dim=D)
    x1, y1, x2, y2 = (t.cuda() for t in (x1, y1, x2, y2))
    f = torch.nn.functional.mse_loss
    sep = []
    for x, y in ((x1, y1), (x2, y2)):
        m.zero_grad(set_to_none=True)
        f(m(x), y).backward()
        sep.append([p.grad.clone() for p in m.parameters()])
    m.zero_grad(set_to_none=True)
    (f(m(x1), y1) + f(m(x2), y2)).backward()
    for p, a, b in zip(m.parameters(), sep[0], sep[1]):
        want = a + b
        assert torch.allclose(p.grad, want, rtol=2e-4, atol=1e-6 * float(want.abs().max()) + 1e-12)
    assert any((a - b).abs().max() > 0 for a, b in zip(sep[0], sep[1]))          # the two passes really differ


@pytest.mark.gpu
def test_trainer_keeps_adam_state_across_a_noop_move():
    """model.to(device) in the middle of training must not reset the fused trainer's Adam moments (torch.optim keeps its state)."""
    from oracle import vis_oracle as V
    from sequoia_pub_b200.train import FusedTrainer
    D, G, B, depth = 1024, 130, 2, 1
    sd = V.make_state_dict(5, G, input_dim=D, depth=depth)
    batches = [V.make_inputs(20 + s, B, G, input_dim=D) for s in range(3)]
    runs = []
    for move in (False, True):
        m = _model(sd, G, D, depth).train()
        tr = FusedTrainer(m, lr=1e-3)
        for s, (x, y) in enumerate(batches):
            if move and s == 2:
                m.to("cuda")                     # nothing moves
                m.float()
            tr.step(x.cuda(), y.cuda())
        runs.append(torch.cat([p.detach().reshape(-1) for p in m.parameters()]).clone())
    assert torch.equal(runs[0], runs[1])
    want = V.train_steps(sd, batches)
    assert abs(tr.loss.item() - want[-1]) / want[-1] < 2e-4
