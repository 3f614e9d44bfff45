"""Drop-in for the model half of the reference's `src/vit.py` (the `--model_type vit` softmax-attention baseline,
SURVEY §8 f-4): `ViT(*, num_outputs, dim, depth, heads, mlp_dim, dim_head=64, num_clusters=100, device='cuda')`
(src/vit.py:93-105) with the reference's module tree and `state_dict()` keys (`pos_emb1D`,
`transformer.layers.{l}.0.{norm,to_qkv,to_out}`, `transformer.layers.{l}.1.net.{0,1,3}`, `linear_head.{0,1}`), a
replaceable `linear_head` (src/main.py:155-157) and the constructor arguments of src/main.py:141-143,161-163,196-198.

`ViT.forward` and its backward are `sq_vit_forward` / `sq_vit_backward` (csrc/vit.cu): split-precision tcgen05 GEMMs for
every Linear and an fp32 shared-memory softmax-attention kernel per (slide, head), behind one `torch.autograd.Function`;
`tformer_lin.FusedAdamW` and `train.FusedTrainer` work with it unchanged.  No CPU / PyTorch fallback: the sub-modules
are parameter containers.  The training-loop half of src/vit.py (`train`, `evaluate`, `predict`) is control plane and
keeps running as it is on top of this class.
"""
import torch
import torch.nn as nn

from . import _lib
from .tformer_lin import _FlatAggregator

_NO_FWD = ("sequoia_b200: only ViT.forward is implemented (one fused CUDA path for the whole aggregator); "
           "sub-modules are parameter containers")


class FeedForward(nn.Module):
    """src/vit.py:39-48."""

    def __init__(self, dim, hidden_dim):
        super().__init__()
        self.net = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, hidden_dim), nn.GELU(), nn.Linear(hidden_dim, dim))

    def forward(self, x):
        raise NotImplementedError(_NO_FWD)


class Attention(nn.Module):
    """src/vit.py:51-61."""

    def __init__(self, dim, heads=8, dim_head=64):
        super().__init__()
        inner_dim = dim_head * heads
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.norm = nn.LayerNorm(dim)
        self.attend = nn.Softmax(dim=-1)
        self.to_qkv = nn.Linear(dim, inner_dim * 3, bias=False)
        self.to_out = nn.Linear(inner_dim, dim, bias=False)

    def forward(self, x):
        raise NotImplementedError(_NO_FWD)


class Transformer(nn.Module):
    """src/vit.py:79-86."""

    def __init__(self, dim, depth, heads, dim_head, mlp_dim):
        super().__init__()
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(nn.ModuleList([Attention(dim, heads=heads, dim_head=dim_head), FeedForward(dim, mlp_dim)]))

    def forward(self, x):
        raise NotImplementedError(_NO_FWD)


class ViT(_FlatAggregator, nn.Module):
    _C = dict(table_len="sq_vit_param_table_len", layout="sq_vit_param_layout", act="sq_vit_act_bytes", bwd_bytes="sq_vit_bwd_bytes",
              fwd="sq_vit_forward", bwd="sq_vit_backward")
    _PER_LAYER = 10
    _NAME = "ViT"

    def __init__(self, *, num_outputs, dim, depth, heads, mlp_dim, dim_head=64, num_clusters=100, device='cuda'):
        super().__init__()
        if dim_head != 64:
            raise NotImplementedError("sequoia_b200 ViT implements dim_head = 64 (the value hard-coded by the reference, src/main.py:143,163)")
        # same construction order as the reference, so the same torch seed gives the same initial weights
        self.pos_emb1D = nn.Parameter(torch.randn(num_clusters, dim))
        self.transformer = Transformer(dim, depth, heads, dim_head, mlp_dim)
        self.to_latent = nn.Identity()
        self.linear_head = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, num_outputs))
        self.device = device
        self._init_flat_state()

    # ------------------------------------------------------------------ layout
    def _config(self):
        head = self.linear_head
        if not (isinstance(head, nn.Sequential) and len(head) == 2 and isinstance(head[0], nn.LayerNorm)
                and isinstance(head[1], nn.Linear)):
            raise RuntimeError("linear_head must be nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, num_outputs))")
        n, d = self.pos_emb1D.shape
        if head[1].in_features != d or head[0].normalized_shape != (d,):
            raise RuntimeError("linear_head does not match dim")
        layers = self.transformer.layers
        attn, ff = layers[0]
        return _lib.VitConfig(d, len(layers), attn.heads, n, head[1].out_features, ff.net[1].out_features)

    def _slots(self, cfg):
        """[(parameter, flat element offset)] for every parameter, from the C layout table."""
        table, total = self._layout_table(cfg)
        slots = [(self.pos_emb1D, table[0])]
        for l, (attn, ff) in enumerate(self.transformer.layers):
            ag, ab, wqkv, wo, fg, fb, w1, b1, w2, b2 = table[1 + 10 * l: 11 + 10 * l]
            slots += [(attn.norm.weight, ag), (attn.norm.bias, ab), (attn.to_qkv.weight, wqkv), (attn.to_out.weight, wo),
                      (ff.net[0].weight, fg), (ff.net[0].bias, fb), (ff.net[1].weight, w1), (ff.net[1].bias, b1),
                      (ff.net[3].weight, w2), (ff.net[3].bias, b2)]
        hg, hb, wh, bh = table[-4:]
        slots += [(self.linear_head[0].weight, hg), (self.linear_head[0].bias, hb), (self.linear_head[1].weight, wh),
                  (self.linear_head[1].bias, bh)]
        return slots, total
