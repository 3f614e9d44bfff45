"""ORACLE (test infrastructure, not product code) — CPU restatement of the reference's softmax-attention ViT baseline.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

Restates, as plain functional PyTorch on the CPU (fp32, or fp64 when the state dict / inputs are double):
  * FeedForward.forward   src/vit.py:47-48 (net = LN, Linear, GELU, Linear :41-46)
  * Attention.forward     src/vit.py:62-74 (LN inside, bias-free to_qkv / to_out, softmax(q k^T * dim_head^-0.5) v)
  * Transformer.forward   src/vit.py:87-91
  * ViT.forward           src/vit.py:107-116
  * one optimisation step src/vit.py:163-166,175-180 (shared with vis_oracle: MSELoss + AdamW)
Pinned against the reference class itself: tests/golden/gen_golden.py imports /root/reference/src/vit.py (with the
tkinter-importing `src.he2rna` stubbed), loads the weights generated here and stores its outputs
(tests/golden/vit_golden.npz); tests/test_oracle_cpu.py checks this restatement against that file.
"""
import torch
import torch.nn.functional as F

from .vis_oracle import AdamW, make_inputs, to_double  # noqa: F401  (same step, same synthetic inputs)


def param_names(depth):
    """state_dict keys in the reference's registration order."""
    names = ["pos_emb1D"]
    for l in range(depth):
        a, q = f"transformer.layers.{l}.0", f"transformer.layers.{l}.1.net"
        names += [f"{a}.norm.weight", f"{a}.norm.bias", f"{a}.to_qkv.weight", f"{a}.to_out.weight",
                  f"{q}.0.weight", f"{q}.0.bias", f"{q}.1.weight", f"{q}.1.bias", f"{q}.3.weight", f"{q}.3.bias"]
    names += ["linear_head.0.weight", "linear_head.0.bias", "linear_head.1.weight", "linear_head.1.bias"]
    return names


def make_state_dict(seed, num_outputs, dim=2048, depth=6, heads=16, mlp_dim=2048, dim_head=64, num_clusters=100):
    """Deterministic weights: nn.Linear-like uniform(+-1/sqrt(fan_in)), NON-trivial LayerNorm affines, randn positional
    embedding (src/vit.py:97).  q/k weights are scaled up 4x so the softmax is far from uniform (a uniform softmax would
    hide errors in the attention path)."""
    g = torch.Generator().manual_seed(seed)
    inner = heads * dim_head

    def lin(out_f, in_f):
        b = 1.0 / in_f ** 0.5
        return (torch.rand(out_f, in_f, generator=g) * 2 - 1) * b, (torch.rand(out_f, generator=g) * 2 - 1) * b

    def ln(n):
        return torch.rand(n, generator=g) + 0.5, torch.randn(n, generator=g) * 0.1

    sd = {"pos_emb1D": torch.randn(num_clusters, dim, generator=g)}
    for l in range(depth):
        a, q = f"transformer.layers.{l}.0", f"transformer.layers.{l}.1.net"
        sd[f"{a}.norm.weight"], sd[f"{a}.norm.bias"] = ln(dim)
        w = lin(3 * inner, dim)[0]
        w[:2 * inner] *= 4.0
        sd[f"{a}.to_qkv.weight"] = w
        sd[f"{a}.to_out.weight"] = lin(dim, inner)[0]
        sd[f"{q}.0.weight"], sd[f"{q}.0.bias"] = ln(dim)
        sd[f"{q}.1.weight"], sd[f"{q}.1.bias"] = lin(mlp_dim, dim)
        sd[f"{q}.3.weight"], sd[f"{q}.3.bias"] = lin(dim, mlp_dim)
    sd["linear_head.0.weight"], sd["linear_head.0.bias"] = ln(dim)
    sd["linear_head.1.weight"], sd["linear_head.1.bias"] = lin(num_outputs, dim)
    assert list(sd.keys()) == param_names(depth)
    return sd


def forward(sd, x, dim_head=64):
    """[B, ..., D] -> [B, num_outputs]."""
    depth = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.layers."))
    B, D = x.shape[0], x.shape[-1]
    x = x.reshape(B, -1, D) + sd["pos_emb1D"]                                              # vit.py:109
    n = x.shape[1]
    for l in range(depth):
        a, q = f"transformer.layers.{l}.0", f"transformer.layers.{l}.1.net"
        h = F.layer_norm(x, (D,), sd[f"{a}.norm.weight"], sd[f"{a}.norm.bias"], 1e-5)      # :63
        qkv = F.linear(h, sd[f"{a}.to_qkv.weight"])                                        # :65
        heads = qkv.shape[-1] // (3 * dim_head)
        qh, kh, vh = (t.reshape(B, n, heads, dim_head).transpose(1, 2) for t in qkv.chunk(3, dim=-1))   # :66
        dots = torch.matmul(qh, kh.transpose(-1, -2)) * dim_head ** -0.5                   # :68
        out = torch.matmul(torch.softmax(dots, dim=-1), vh)                                # :70-72
        out = out.transpose(1, 2).reshape(B, n, heads * dim_head)                          # :73
        x = F.linear(out, sd[f"{a}.to_out.weight"]) + x                                    # :74,89
        h = F.layer_norm(x, (D,), sd[f"{q}.0.weight"], sd[f"{q}.0.bias"], 1e-5)
        h = F.gelu(F.linear(h, sd[f"{q}.1.weight"], sd[f"{q}.1.bias"]))
        x = F.linear(h, sd[f"{q}.3.weight"], sd[f"{q}.3.bias"]) + x                        # :90
    x = x.mean(dim=1)                                                                      # :112
    x = F.layer_norm(x, (D,), sd["linear_head.0.weight"], sd["linear_head.0.bias"], 1e-5)
    return F.linear(x, sd["linear_head.1.weight"], sd["linear_head.1.bias"])               # :115


def loss_and_grads(sd, x, y):
    """MSELoss (mean over B*G, src/vit.py:129,166) and its gradient w.r.t. every parameter (src/vit.py:179)."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    pred = forward(params, x)
    loss = F.mse_loss(pred, y)
    grads = torch.autograd.grad(loss, list(params.values()))
    return loss.detach(), pred.detach(), dict(zip(params.keys(), grads))


def train_steps(sd, batches, lr=1e-3):
    """Runs len(batches) optimisation steps in place on sd; returns the per-step losses."""
    opt = AdamW(sd, lr=lr)
    losses = []
    for x, y in batches:
        loss, _, grads = loss_and_grads(sd, x, y)
        opt.step(sd, grads)
        losses.append(float(loss))
    return losses
