"""Drop-in for the reference's `src/tformer_lin.py` (ViS = SummaryMixing "linearized attention" aggregator).

Same classes, constructor signatures, module tree and `state_dict()` keys as the reference
(`ViS(num_outputs, input_dim, depth, nheads, dimensions_f, dimensions_s, dimensions_c, num_clusters=100,
device='cuda:0')`, src/tformer_lin.py:80-95; `PyTorchModelHubMixin` for `from_pretrained`, :4,80;
`linear_head` stays a replaceable `nn.Sequential(LayerNorm, Linear)`, src/main.py:155-157), so `src/main.py`,
`src/vit.py:train/evaluate/predict` and `evaluation/predict_independent_dataset.py` can use it unchanged.

What differs is where the arithmetic runs: `ViS.forward` and its backward are `sq_vis_forward` / `sq_vis_backward`
(csrc/vis.cu) — split-precision tcgen05 GEMMs with fused epilogues on a flat parameter buffer — exposed to autograd as
one `torch.autograd.Function`, so `loss.backward()` + any `torch.optim` optimizer keep working; `FusedAdamW` below is an
opt-in `torch.optim.Optimizer` that updates the flat buffer (and the bf16 weight planes) in one kernel.
There is no PyTorch/CPU fallback: the sub-modules are parameter containers and only `ViS.forward` computes.
"""
import ctypes as C
import weakref

import torch
import torch.nn as nn

from . import _lib

try:
    from huggingface_hub import PyTorchModelHubMixin
except Exception:  # pragma: no cover - huggingface_hub is an optional dependency of the reference too
    class PyTorchModelHubMixin:  # type: ignore
        pass

_NO_FWD = ("sequoia_b200: only ViS.forward is implemented (one fused CUDA path for the whole aggregator); "
           "sub-modules are parameter containers")


class SummaryMixing(nn.Module):
    """Parameters of one head (src/tformer_lin.py:8-16)."""

    def __init__(self, input_dim, dimensions_f, dimensions_s, dimensions_c):
        super().__init__()
        self.local_norm = nn.LayerNorm(dimensions_f)
        self.summary_norm = nn.LayerNorm(dimensions_s)
        self.s = nn.Linear(input_dim, dimensions_s)
        self.f = nn.Linear(input_dim, dimensions_f)
        self.c = nn.Linear(dimensions_s + dimensions_f, dimensions_c)

    def forward(self, x):
        raise NotImplementedError(_NO_FWD)


class MultiHeadSummary(nn.Module):
    """src/tformer_lin.py:29-37."""

    def __init__(self, nheads, input_dim, dimensions_f, dimensions_s, dimensions_c, dimensions_projection):
        super().__init__()
        self.mixers = nn.ModuleList([])
        for _ in range(nheads):
            self.mixers.append(SummaryMixing(input_dim=input_dim, dimensions_f=dimensions_f, dimensions_s=dimensions_s,
                                             dimensions_c=dimensions_c))
        self.projection = nn.Linear(nheads * dimensions_c, dimensions_projection)

    def forward(self, x):
        raise NotImplementedError(_NO_FWD)


class FeedForward(nn.Module):
    """src/tformer_lin.py:51-59."""

    def __init__(self, dim, hidden_dim):
        super().__init__()
        self.net = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, hidden_dim), nn.GELU(), nn.Linear(hidden_dim, dim))

    def forward(self, x):
        raise NotImplementedError(_NO_FWD)


class SummaryTransformer(nn.Module):
    """src/tformer_lin.py:64-72."""

    def __init__(self, input_dim, depth, nheads, dimensions_f, dimensions_s, dimensions_c):
        super().__init__()
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(nn.ModuleList([
                MultiHeadSummary(nheads, input_dim, dimensions_f, dimensions_s, dimensions_c, dimensions_projection=input_dim),
                FeedForward(input_dim, input_dim)]))

    def forward(self, x):
        raise NotImplementedError(_NO_FWD)


class _ViSFunction(torch.autograd.Function):
    """pred = ViS(x); all parameters are inputs so autograd delivers their gradients the usual way."""

    @staticmethod
    def forward(ctx, model, x, *params):
        pred, act = model._forward_impl(x, keep=True)
        ctx.model, ctx.act, ctx.batch, ctx.token = model, act, x.shape[0], model._flat_token
        ctx.need_dx = x.requires_grad
        ctx.x_shape = x.shape
        return pred

    @staticmethod
    def backward(ctx, dpred):
        model = ctx.model
        if ctx.token != model._flat_token:
            raise RuntimeError("sequoia_b200: parameters were re-laid out between forward and backward")
        gflat, dx = model._backward_impl(ctx.act, dpred, ctx.batch, ctx.need_dx)
        ctx.act = None
        grads = model._split_views(gflat)
        return (None, dx.view(ctx.x_shape) if dx is not None else None, *grads)


class _FlatAggregator:
    """Host side shared by the two aggregators (`ViS` here, `ViT` in vit.py): the nn.Parameters are views of ONE flat fp32
    buffer laid out by the C library, forward/backward are one C call each (stage by stage for the trainer), gradients come
    back as views of a flat gradient buffer.  Subclasses provide `_config()`, `_slots(cfg)`, the names of their C entry
    points (`_C`) and the number of layout-table entries per layer (`_PER_LAYER`)."""

    _C = {}
    _PER_LAYER = 0
    _NAME = "model"

    def _init_flat_state(self):
        self._flat = None            # fp32 flat parameter buffer the nn.Parameters are views of
        self._flat_token = 0         # bumped on every re-layout
        self._w_hi = self._w_lo = None
        self._planes_key = None      # parameter-version fingerprint the planes were computed from
        self._gbufs = [None, None]   # flat gradient buffers (alternated so accumulation into .grad stays correct)
        self._act_cache = None
        self._scratch = None

    def _layout_table(self, cfg):
        L = _lib.lib()
        n = getattr(L, self._C["table_len"])(C.byref(cfg))
        if n < 0:
            _lib.check(n)
        table = (C.c_longlong * n)()
        total = C.c_longlong()
        _lib.check(getattr(L, self._C["layout"])(C.byref(cfg), table, n, C.byref(total)))
        return list(table), total.value

    def _stage_ranges(self):
        """[begin, end) element ranges of the flat buffer per backward stage: l < depth = layer l (stage 0 also holds
        pos_emb1D), index depth = regression head."""
        table, total = self._layout_table(self._cfg)
        depth, k = self._cfg.depth, self._PER_LAYER
        starts = [table[1 + k * l] for l in range(depth)] + [table[len(table) - 4], total]
        return [(0 if l == 0 else starts[l], starts[l + 1]) for l in range(depth)] + [(starts[depth], total)]

    # ------------------------------------------------------------------ layout
    def __setattr__(self, name, value):
        super().__setattr__(name, value)
        if name in ("linear_head", "pos_emb1D", "transformer") and "_flat" in self.__dict__:
            self.__dict__["_flat"] = None      # e.g. main.py:155-157 replaces the head after loading a checkpoint

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        # .to()/.cuda()/.float() re-allocate the parameters only when something actually changes: keep the flat buffer (and with
        # it the optimizer state bound to it) when every parameter still is the view it was
        flat = self.__dict__.get("_flat")
        if flat is not None:
            base = flat.data_ptr()
            try:
                slots, _ = self._slots(self._cfg)
                same = all(p.data_ptr() == base + 4 * off and p.dtype == torch.float32 for p, off in slots)
            except Exception:
                same = False
            if not same:
                self.__dict__["_flat"] = None
        return out

    def _ensure_flat(self):
        p0, p1 = self.pos_emb1D, self.linear_head[1].bias
        if self._flat is not None:
            base = self._flat.data_ptr()
            if (p0.data_ptr() == base + 4 * self._off0 and p1.data_ptr() == base + 4 * self._off1
                    and self.linear_head[1].out_features == self._cfg.num_outputs):
                return
        dev = p0.device
        if dev.type != "cuda":
            raise RuntimeError(f"sequoia_b200 {self._NAME} runs on a B200 only: call .to('cuda') first (no CPU fallback)")
        _lib.require_device()
        cfg = self._config()
        slots, total = self._slots(cfg)
        for p, _ in slots:
            if p.dtype != torch.float32 or p.device != dev:
                raise RuntimeError(f"{self._NAME} parameters must all be float32 on one CUDA device")
        flat = torch.zeros(total, dtype=torch.float32, device=dev)
        sizes, order = [], []
        pos = 0
        for p, off in slots:                         # offsets ascend in slot order except per-head interleaving: sort
            order.append((off, p))
        order.sort(key=lambda t: t[0])
        with torch.no_grad():
            for off, p in order:
                n = p.numel()
                flat[off:off + n].copy_(p.detach().reshape(-1))
                p.data = flat[off:off + n].view(p.shape)
                p._sq_vis_owner = weakref.ref(self)
                if off > pos:
                    sizes.append(off - pos)
                sizes.append(n)
                pos = off + n
        if total > pos:
            sizes.append(total - pos)
        # bookkeeping to hand autograd per-parameter views of a flat gradient buffer in parameter order
        chunk_of, idx, pos = {}, 0, 0
        for off, p in order:
            if off > pos:
                idx += 1
            chunk_of[id(p)] = idx
            idx += 1
            pos = off + p.numel()
        self._split_sizes = sizes
        self._params = [p for p, _ in slots]
        self._param_chunks = [(chunk_of[id(p)], tuple(p.shape)) for p in self._params]
        self._flat, self._cfg, self._total = flat, cfg, total
        self._off0, self._off1 = slots[0][1], slots[-1][1]
        self._w_hi = torch.empty(total, dtype=torch.bfloat16, device=dev)
        self._w_lo = torch.empty(total, dtype=torch.bfloat16, device=dev)
        self._planes_key = None
        self._gbufs = [None, None]
        self._act_cache = None
        self._scratch = None
        self._flat_token += 1

    def _split_views(self, flat):
        chunks = flat.split_with_sizes(self._split_sizes)
        return [chunks[i].view(shape) for i, shape in self._param_chunks]

    def _version_key(self):
        return sum(p._version for p in self._params)

    @_lib.with_device_of(lambda self: self._flat)
    def _refresh_planes(self):
        key = self._version_key()
        if key != self._planes_key:
            _lib.check(_lib.lib().sq_split_bf16(_lib.ptr(self._flat), _lib.ptr(self._w_hi), _lib.ptr(self._w_lo), 1, self._total,
                                                self._total, self._total, _lib.stream_ptr()))
            self._planes_key = key

    def _planes_are_fresh(self):
        """Called by the fused optimizer after it rewrote params and planes in one kernel."""
        self._planes_key = self._version_key()

    # ------------------------------------------------------------------ compute
    @_lib.with_device_of(lambda self, x, *a, **k: x)
    def _forward_impl(self, x, keep):
        self._ensure_flat()
        cfg = self._cfg
        if x.device != self._flat.device or x.dtype != torch.float32:
            raise ValueError(f"{self._NAME}.forward expects a float32 tensor on the model's CUDA device")
        B = x.shape[0]
        if B == 0:
            return torch.empty(0, cfg.num_outputs, dtype=torch.float32, device=x.device), None
        x = x.reshape(B, -1, x.shape[-1]).contiguous()            # rearrange 'b ... d -> b (...) d' (tformer_lin.py:100)
        if x.shape[1] != cfg.num_clusters or x.shape[2] != cfg.input_dim:
            raise ValueError(f"expected [B, {cfg.num_clusters}, {cfg.input_dim}] cluster features, got {tuple(x.shape)}")
        self._refresh_planes()
        L = _lib.lib()
        need = getattr(L, self._C["act"])(C.byref(cfg), B)
        if keep:
            act = torch.empty(need, dtype=torch.uint8, device=x.device)
        else:
            if self._act_cache is None or self._act_cache.numel() < need:
                self._act_cache = torch.empty(need, dtype=torch.uint8, device=x.device)
            act = self._act_cache
        pred = torch.empty(B, cfg.num_outputs, dtype=torch.float32, device=x.device)
        if B > 0:
            _lib.check(getattr(L, self._C["fwd"])(C.byref(cfg), _lib.ptr(self._flat), _lib.ptr(self._w_hi), _lib.ptr(self._w_lo), _lib.ptr(x), B,
                                        _lib.ptr(pred), _lib.ptr(act), act.numel(), _lib.stream_ptr()))
        return pred, act

    def _grad_buffer(self):
        """A flat gradient buffer nobody else references.  The per-parameter gradients handed to autograd are VIEWS of the buffer:
        AccumulateGrad may keep them as `.grad`, and when the model is called several times inside one graph
        (`loss = f(model(x1)) + f(model(x2))`) autograd's input buffers hold the views of the first backward while the second one
        runs.  A buffer is reused only when its storage has no other user (tensor + the probing storage handle = 2 references)."""
        for b in self._gbufs:
            if b is not None and torch._C._storage_Use_Count(b.untyped_storage()._cdata) <= 2:
                return b
        b = torch.zeros(self._total, dtype=torch.float32, device=self._flat.device)
        self._gbufs = [g for g in self._gbufs if g is not None] + [b]
        return b

    def _scratch_for(self, B):
        need = getattr(_lib.lib(), self._C["bwd_bytes"])(C.byref(self._cfg), B)
        if self._scratch is None or self._scratch.numel() < need:
            self._scratch = torch.empty(need, dtype=torch.uint8, device=self._flat.device)
        return self._scratch

    @_lib.with_device_of(lambda self, act, *a, **k: act if act is not None else self._flat)
    def _backward_impl(self, act, dpred, B, need_dx, gbuf=None, stage_hi=None, stage_lo=0):
        cfg = self._cfg
        if gbuf is None:
            gbuf = self._grad_buffer()
        scratch = self._scratch_for(B)
        dx = torch.empty(B, cfg.num_clusters, cfg.input_dim, dtype=torch.float32, device=gbuf.device) if need_dx else None
        if dpred is not None:
            dpred = dpred.contiguous()
        if B > 0:
            _lib.check(getattr(_lib.lib(), self._C["bwd"])(C.byref(cfg), _lib.ptr(self._flat), _lib.ptr(self._w_hi), _lib.ptr(self._w_lo),
                                                  _lib.ptr(dpred), B, _lib.ptr(act), act.numel(), _lib.ptr(gbuf), _lib.ptr(dx),
                                                  _lib.ptr(scratch), scratch.numel(), cfg.depth if stage_hi is None else stage_hi,
                                                  stage_lo, _lib.stream_ptr()))
        else:
            gbuf.zero_()
        return gbuf, dx

    def forward(self, x):
        self._ensure_flat()
        n = self._cfg.num_clusters
        if x.dim() >= 2 and n > 1 and x.numel() == x.shape[0] * x.shape[-1]:
            # one token per slide: `x + pos_emb1D` broadcasts it over the N positions in the reference (tformer_lin.py:100,
            # vit.py:109) — which is what spatial_vis/visualize.py:80 relies on when it passes an unbatched [100, D] tensor
            x = x.reshape(x.shape[0], 1, x.shape[-1]).expand(x.shape[0], n, x.shape[-1])
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self._params)):
            return _ViSFunction.apply(self, x, *self._params)
        return self._forward_impl(x, keep=False)[0]


class ViS(_FlatAggregator, nn.Module, PyTorchModelHubMixin):
    _C = dict(table_len="sq_vis_param_table_len", layout="sq_vis_param_layout", act="sq_vis_act_bytes", bwd_bytes="sq_vis_bwd_bytes",
              fwd="sq_vis_forward", bwd="sq_vis_backward")
    _PER_LAYER = 18
    _NAME = "ViS"

    def __init__(self, num_outputs, input_dim, depth, nheads, dimensions_f, dimensions_s, dimensions_c, num_clusters=100,
                 device='cuda:0'):
        super().__init__()
        if not (dimensions_f == dimensions_s == dimensions_c == 64):
            raise NotImplementedError("sequoia_b200 ViS implements dimensions_f = dimensions_s = dimensions_c = 64 "
                                      "(the values hard-coded by the reference, src/main.py:147,167)")
        # same construction order as the reference, so the same torch seed gives the same initial weights
        self.pos_emb1D = nn.Parameter(torch.randn(num_clusters, input_dim))
        self.transformer = SummaryTransformer(input_dim, depth, nheads, dimensions_f, dimensions_s, dimensions_c)
        self.to_latent = nn.Identity()
        self.linear_head = nn.Sequential(nn.LayerNorm(input_dim), nn.Linear(input_dim, num_outputs))
        self.device = device
        self._init_flat_state()

    # ------------------------------------------------------------------ layout
    def _config(self):
        head = self.linear_head
        if not (isinstance(head, nn.Sequential) and len(head) == 2 and isinstance(head[0], nn.LayerNorm)
                and isinstance(head[1], nn.Linear)):
            raise RuntimeError("linear_head must be nn.Sequential(nn.LayerNorm(D), nn.Linear(D, num_outputs))")
        n, d = self.pos_emb1D.shape
        if head[1].in_features != d or head[0].normalized_shape != (d,):
            raise RuntimeError("linear_head does not match input_dim")
        layers = self.transformer.layers
        return _lib.VisConfig(d, len(layers), len(layers[0][0].mixers), n, head[1].out_features)

    def _slots(self, cfg):
        """[(parameter, flat element offset)] for every parameter, from the C layout table."""
        table, total = self._layout_table(cfg)
        n = len(table)
        D = cfg.input_dim
        slots = [(self.pos_emb1D, table[0])]
        for l, (attn, ff) in enumerate(self.transformer.layers):
            lnl_g, lnl_b, lns_g, lns_b, ws, bs, wf, bf, wc, bc, wp, bp, fg, fb, w1, b1, w2, b2 = table[1 + 18 * l: 19 + 18 * l]
            for h, m in enumerate(attn.mixers):
                slots += [(m.local_norm.weight, lnl_g + 64 * h), (m.local_norm.bias, lnl_b + 64 * h),
                          (m.summary_norm.weight, lns_g + 64 * h), (m.summary_norm.bias, lns_b + 64 * h),
                          (m.s.weight, ws + 64 * D * h), (m.s.bias, bs + 64 * h),
                          (m.f.weight, wf + 64 * D * h), (m.f.bias, bf + 64 * h),
                          (m.c.weight, wc + 64 * 128 * h), (m.c.bias, bc + 64 * h)]
            slots += [(attn.projection.weight, wp), (attn.projection.bias, bp), (ff.net[0].weight, fg), (ff.net[0].bias, fb),
                      (ff.net[1].weight, w1), (ff.net[1].bias, b1), (ff.net[3].weight, w2), (ff.net[3].bias, b2)]
        hg, hb, wh, bh = table[n - 4: n]
        slots += [(self.linear_head[0].weight, hg), (self.linear_head[0].bias, hb), (self.linear_head[1].weight, wh),
                  (self.linear_head[1].bias, bh)]
        return slots, total


class FusedAdamW(torch.optim.Optimizer):
    """`torch.optim.AdamW(params, lr, betas, eps, weight_decay, amsgrad=False)` for the parameters of ONE `ViS` model,
    executed as a single kernel over the model's flat parameter buffer (csrc/vis.cu: adamw_kernel), which also rewrites
    the bf16 weight planes the next forward needs.  Same constructor / step() / zero_grad() surface (src/main.py:180-183,
    src/vit.py:175-180)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False, grad_scale=1.0):
        if amsgrad:
            raise NotImplementedError("FusedAdamW: amsgrad=False only (the reference's setting)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) != 1:
            raise NotImplementedError("FusedAdamW: one parameter group")
        self.grad_scale = grad_scale
        self._model_ref = None
        self._token = None
        self._m = self._v = None
        self._step = 0
        self._step_t = torch.tensor(0.0)          # shared by every per-parameter state entry (torch keeps one per tensor)

    def _bind(self):
        ps = self.param_groups[0]["params"]
        owner = getattr(ps[0], "_sq_vis_owner", None)
        model = owner() if owner is not None else None
        if model is None:
            raise RuntimeError("FusedAdamW: parameters do not belong to a sequoia_b200 ViS that has run a forward pass")
        model._ensure_flat()
        if {id(p) for p in ps} != {id(p) for p in model._params}:
            raise RuntimeError("FusedAdamW must own exactly the parameters of one ViS model")
        if self._token != model._flat_token:
            old_m, old_v, old_params = self._m, self._v, getattr(self, "_bound_params", None)
            self._m = torch.zeros_like(model._flat)
            self._v = torch.zeros_like(model._flat)
            if old_m is not None:          # re-layout (e.g. new head): carry over the moments of surviving parameters
                old_views = dict(zip(map(id, old_params), zip(self._views(old_m, self._old_split, self._old_chunks),
                                                              self._views(old_v, self._old_split, self._old_chunks))))
                for p, mv, vv in zip(model._params, model._split_views(self._m), model._split_views(self._v)):
                    if id(p) in old_views and old_views[id(p)][0].shape == mv.shape:
                        mv.copy_(old_views[id(p)][0]); vv.copy_(old_views[id(p)][1])
            self._bound_params, self._old_split, self._old_chunks = list(model._params), model._split_sizes, model._param_chunks
            self._token = model._flat_token
            for p, mv, vv in zip(model._params, model._split_views(self._m), model._split_views(self._v)):
                self.state[p] = {"step": self._step_t, "exp_avg": mv, "exp_avg_sq": vv}
        self._model_ref = weakref.ref(model)
        return model

    @staticmethod
    def _views(flat, split, chunks):
        parts = flat.split_with_sizes(split)
        return [parts[i].view(shape) for i, shape in chunks]

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        model = self._bind()
        g0 = model.pos_emb1D.grad
        if g0 is None:
            return loss
        gbuf = None
        for b in model._gbufs:
            if b is not None and g0.data_ptr() == b.data_ptr() + 4 * model._off0:
                gbuf = b
        if gbuf is None or model.linear_head[1].bias.grad is None or \
                model.linear_head[1].bias.grad.data_ptr() != gbuf.data_ptr() + 4 * model._off1:
            # gradients live elsewhere (cloned by autograd, or produced by another graph): gather them into a flat buffer
            gbuf = model._grad_buffer()
            views = model._split_views(gbuf)
            torch._foreach_copy_(views, [p.grad if p.grad is not None else torch.zeros_like(p) for p in model._params])
        grp = self.param_groups[0]
        self._step += 1
        self._step_t.fill_(float(self._step))
        _lib.check(_lib.lib().sq_adamw_flat(_lib.ptr(model._flat), _lib.ptr(gbuf), _lib.ptr(self._m), _lib.ptr(self._v),
                                            _lib.ptr(model._w_hi), _lib.ptr(model._w_lo), model._total, grp["lr"], grp["betas"][0],
                                            grp["betas"][1], grp["eps"], grp["weight_decay"], self._step, self.grad_scale,
                                            _lib.stream_ptr()))
        model._planes_are_fresh()
        return loss
