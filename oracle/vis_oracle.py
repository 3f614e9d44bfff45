"""ORACLE (test infrastructure, not product code) — CPU restatement of the reference ViS aggregator.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

Restates, as plain functional PyTorch on the CPU (fp32, or fp64 when the state dict / inputs are double):
  * SummaryMixing.forward        src/tformer_lin.py:18-26
  * MultiHeadSummary.forward     src/tformer_lin.py:39-48
  * FeedForward.forward          src/tformer_lin.py:60-61 (net = LN, Linear, GELU, Linear :54-59)
  * SummaryTransformer.forward   src/tformer_lin.py:73-77
  * ViS.forward                  src/tformer_lin.py:97-106
  * one optimisation step        src/vit.py:163-166,175-180 with AdamW(lr, weight_decay=0, amsgrad=False) src/main.py:180-183
Pinned against the reference class itself: tests/golden/gen_golden.py imports /root/reference/src/tformer_lin.py,
loads the weights generated here and stores its outputs (tests/golden/vis_golden.npz);
tests/test_oracle_cpu.py checks this restatement against that file.
"""
import torch
import torch.nn.functional as F


def param_names(depth, nheads):
    """state_dict keys in the reference's registration order."""
    names = ["pos_emb1D"]
    for l in range(depth):
        for h in range(nheads):
            p = f"transformer.layers.{l}.0.mixers.{h}"
            names += [f"{p}.local_norm.weight", f"{p}.local_norm.bias", f"{p}.summary_norm.weight", f"{p}.summary_norm.bias",
                      f"{p}.s.weight", f"{p}.s.bias", f"{p}.f.weight", f"{p}.f.bias", f"{p}.c.weight", f"{p}.c.bias"]
        names += [f"transformer.layers.{l}.0.projection.weight", f"transformer.layers.{l}.0.projection.bias"]
        q = f"transformer.layers.{l}.1.net"
        names += [f"{q}.0.weight", f"{q}.0.bias", f"{q}.1.weight", f"{q}.1.bias", f"{q}.3.weight", f"{q}.3.bias"]
    names += ["linear_head.0.weight", "linear_head.0.bias", "linear_head.1.weight", "linear_head.1.bias"]
    return names


def make_state_dict(seed, num_outputs, input_dim=2048, depth=6, nheads=16, d=64, num_clusters=100):
    """Deterministic weights: nn.Linear-like uniform(+-1/sqrt(fan_in)) matrices and biases, NON-trivial LayerNorm
    affines (default 1/0 would hide LN bugs), randn positional embedding (src/tformer_lin.py:86)."""
    g = torch.Generator().manual_seed(seed)
    D = input_dim

    def lin(out_f, in_f):
        b = 1.0 / in_f ** 0.5
        return (torch.rand(out_f, in_f, generator=g) * 2 - 1) * b, (torch.rand(out_f, generator=g) * 2 - 1) * b

    def ln(n):
        return torch.rand(n, generator=g) + 0.5, torch.randn(n, generator=g) * 0.1

    sd = {"pos_emb1D": torch.randn(num_clusters, D, generator=g)}
    for l in range(depth):
        for h in range(nheads):
            p = f"transformer.layers.{l}.0.mixers.{h}"
            sd[f"{p}.local_norm.weight"], sd[f"{p}.local_norm.bias"] = ln(d)
            sd[f"{p}.summary_norm.weight"], sd[f"{p}.summary_norm.bias"] = ln(d)
            sd[f"{p}.s.weight"], sd[f"{p}.s.bias"] = lin(d, D)
            sd[f"{p}.f.weight"], sd[f"{p}.f.bias"] = lin(d, D)
            sd[f"{p}.c.weight"], sd[f"{p}.c.bias"] = lin(d, 2 * d)
        sd[f"transformer.layers.{l}.0.projection.weight"], sd[f"transformer.layers.{l}.0.projection.bias"] = lin(D, nheads * d)
        q = f"transformer.layers.{l}.1.net"
        sd[f"{q}.0.weight"], sd[f"{q}.0.bias"] = ln(D)
        sd[f"{q}.1.weight"], sd[f"{q}.1.bias"] = lin(D, D)
        sd[f"{q}.3.weight"], sd[f"{q}.3.bias"] = lin(D, D)
    sd["linear_head.0.weight"], sd["linear_head.0.bias"] = ln(D)
    sd["linear_head.1.weight"], sd["linear_head.1.bias"] = lin(num_outputs, D)
    assert list(sd.keys()) == param_names(depth, nheads)
    return sd


def make_inputs(seed, batch, num_outputs, input_dim=2048, num_clusters=100):
    """x = relu(randn)*0.5 (cluster means of post-ReLU features are non-negative), y = rand*10 (log-FPKM-like range)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.relu(torch.randn(batch, num_clusters, input_dim, generator=g)) * 0.5
    y = torch.rand(batch, num_outputs, generator=g) * 10
    return x, y


def _depth_heads(sd):
    depth = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.layers."))
    nheads = 1 + max(int(k.split(".")[5]) for k in sd if ".mixers." in k)
    return depth, nheads


def forward(sd, x):
    """[B, ..., D] -> [B, num_outputs]."""
    depth, nheads = _depth_heads(sd)
    B, D = x.shape[0], x.shape[-1]
    x = x.reshape(B, -1, D) + sd["pos_emb1D"]                                  # tformer_lin.py:100
    for l in range(depth):
        outs = []
        for h in range(nheads):                                                # :42-43
            p = f"transformer.layers.{l}.0.mixers.{h}"
            d = sd[f"{p}.f.bias"].shape[0]
            local = F.gelu(F.layer_norm(F.linear(x, sd[f"{p}.f.weight"], sd[f"{p}.f.bias"]), (d,),
                                        sd[f"{p}.local_norm.weight"], sd[f"{p}.local_norm.bias"], 1e-5))   # :20
            t = F.linear(x, sd[f"{p}.s.weight"], sd[f"{p}.s.bias"])                                        # :21
            t = F.gelu(F.layer_norm(t.mean(dim=1), (t.shape[-1],), sd[f"{p}.summary_norm.weight"],
                                    sd[f"{p}.summary_norm.bias"], 1e-5))                                   # :22
            t = t.unsqueeze(1).repeat(1, x.shape[1], 1)                                                    # :23
            outs.append(F.gelu(F.linear(torch.cat([local, t], dim=-1), sd[f"{p}.c.weight"], sd[f"{p}.c.bias"])))  # :24
        a = f"transformer.layers.{l}.0.projection"
        x = F.linear(torch.cat(outs, dim=-1), sd[f"{a}.weight"], sd[f"{a}.bias"]) + x                     # :45-46,75
        q = f"transformer.layers.{l}.1.net"
        hdn = F.layer_norm(x, (D,), sd[f"{q}.0.weight"], sd[f"{q}.0.bias"], 1e-5)
        hdn = F.gelu(F.linear(hdn, sd[f"{q}.1.weight"], sd[f"{q}.1.bias"]))
        x = F.linear(hdn, sd[f"{q}.3.weight"], sd[f"{q}.3.bias"]) + x                                      # :76
    x = x.mean(dim=1)                                                                                      # :103
    x = F.layer_norm(x, (D,), sd["linear_head.0.weight"], sd["linear_head.0.bias"], 1e-5)
    return F.linear(x, sd["linear_head.1.weight"], sd["linear_head.1.bias"])                               # :106


def loss_and_grads(sd, x, y):
    """MSELoss (mean over B*G, src/vit.py:129,166) and its gradient w.r.t. every parameter (src/vit.py:179)."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    pred = forward(params, x)
    loss = F.mse_loss(pred, y)
    grads = torch.autograd.grad(loss, list(params.values()))
    return loss.detach(), pred.detach(), dict(zip(params.keys(), grads))


class AdamW:
    """torch.optim.AdamW(lr, betas=(0.9,0.999), eps=1e-8, weight_decay, amsgrad=False) restated (src/main.py:180-183)."""

    def __init__(self, sd, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.lr, self.b1, self.b2, self.eps, self.wd, self.t = lr, betas[0], betas[1], eps, weight_decay, 0
        self.m = {k: torch.zeros_like(v) for k, v in sd.items()}
        self.v = {k: torch.zeros_like(v) for k, v in sd.items()}

    def step(self, sd, grads):
        self.t += 1
        bc1, bc2 = 1 - self.b1 ** self.t, 1 - self.b2 ** self.t
        for k, p in sd.items():
            g = grads[k]
            p.mul_(1 - self.lr * self.wd)
            self.m[k].lerp_(g, 1 - self.b1)
            self.v[k].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            denom = (self.v[k].sqrt() / bc2 ** 0.5).add_(self.eps)
            p.addcdiv_(self.m[k], denom, value=-self.lr / bc1)


def train_steps(sd, batches, lr=1e-3):
    """Runs len(batches) optimisation steps in place on sd; returns the per-step losses."""
    opt = AdamW(sd, lr=lr)
    losses = []
    for x, y in batches:
        loss, _, grads = loss_and_grads(sd, x, y)
        opt.step(sd, grads)
        losses.append(float(loss))
    return losses


def to_double(sd):
    return {k: v.double() for k, v in sd.items()}
