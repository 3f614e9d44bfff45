"""Fused optimisation step of an aggregator (`tformer_lin.ViS` or `vit.ViT`): the body of the reference's training loop (src/vit.py:163-166,175-180 —
`pred = model(x); loss = MSELoss(pred, y); optimizer.zero_grad(); loss.backward(); optimizer.step()` with
`AdamW(lr, weight_decay=0, amsgrad=False)`, src/main.py:180-183) enqueued as one chain of CUDA kernels with no autograd
graph: sq_vis_forward -> sq_mse_fwd_bwd -> sq_vis_backward (stage by stage) -> [NCCL all-reduce] -> sq_adamw_flat.

Data parallel: one process per GPU, the batch of slides is split evenly across ranks, every backward stage's
contiguous slice of the flat gradient buffer is all-reduced (sum) as soon as the stage is enqueued so the exchange of
the head / upper layers overlaps the backward of the lower layers; AdamW then applies grad * 1/world_size, which
equals the gradient of the global-batch mean loss because MSELoss is a mean over equal shards (src/vit.py:129).
"""
import ctypes as C

import torch

from . import _lib
from .dist import allreduce_stage


class FusedTrainer:
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, process_group=None, overlap=True, comm=None,
                 comm_ctas=4):
        """comm: how the data-parallel gradient exchange runs - "multimem" (the library's own NVSwitch-multicast all-reduce over a
        symmetric gradient buffer, `comm_ctas` CTAs, with the persistent GEMMs of the backward pass limited to the remaining SMs),
        "nccl" (torch.distributed all-reduce per stage) or "auto" (multimem when the system has multicast support, else NCCL);
        None reads SQ_DP_COMM (default "auto").  Measured on 8xB200 (profiles/r02_dp_allreduce_study.json): 7.27 ms/step with
        multimem and 4 CTAs (761 GB/s bus bandwidth on the 525 MB buffer) against 7.86 ms with NCCL (699 GB/s)."""
        self.model = model
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.step_count = 0
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        self.overlap = overlap
        import os
        self.comm = (comm or os.environ.get("SQ_DP_COMM", "auto")).lower()
        if self.comm not in ("auto", "multimem", "nccl"):
            raise ValueError("comm must be 'auto', 'multimem' or 'nccl'")
        self.comm_ctas = int(os.environ.get("SQ_DP_COMM_CTAS", comm_ctas))
        self._mm = None
        self._comm_stream = None
        self._token = None
        self._opt_stream = None
        self.launch_count = 0     # sq_* calls of the last step (bench.py reports kernels separately)

    def _bind(self):
        m = self.model
        m._ensure_flat()
        if self._token != m._flat_token:
            dev = m._flat.device
            old = (self.m, self.v, self.stage_range) if self._token is not None else None
            self.m = torch.zeros_like(m._flat)
            self.v = torch.zeros_like(m._flat)
            self.g = torch.zeros_like(m._flat)
            self._mm = None
            if self.world > 1 and self.comm in ("multimem", "auto") and self.overlap:
                from .dist import MultimemAllReduce
                try:
                    self._mm = MultimemAllReduce(m._flat.numel(), dev, self.pg, self.comm_ctas)
                    self.g = self._mm.buf
                except Exception as e:                   # no multicast / symmetric memory on this system
                    if self.comm == "multimem":
                        raise
                    import warnings
                    warnings.warn(f"sequoia_b200: NVSwitch multicast all-reduce unavailable ({e}); the gradient exchange uses NCCL")
                    self.comm = "nccl"
            self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
            self.mse_scratch = torch.empty(1024, dtype=torch.float32, device=dev)
            self._token = m._flat_token
            self.stage_range = m._stage_ranges()         # contiguous slices of the flat buffer per backward stage
            if old is not None:
                # the parameters were re-laid out (a .to() that moved them, a replaced head): the Adam moments follow them, like
                # torch.optim state does - everything when the layout is unchanged, the layers below a replaced head otherwise
                om, ov, orange = old
                if om.numel() == self.m.numel() and orange == self.stage_range:
                    self.m.copy_(om); self.v.copy_(ov)
                elif orange[:-1] == self.stage_range[:-1]:
                    keep = self.stage_range[-1][0]
                    self.m[:keep].copy_(om[:keep]); self.v[:keep].copy_(ov[:keep])
                else:
                    self.step_count = 0                  # a different architecture: nothing carries over, bias correction restarts
        return m

    @_lib.with_device_of(lambda self, x, *a, **k: x)
    def step(self, x, y):
        """x: float32 [B, N, D], y: float32 [B, G] on the model's device -> loss (1-element device tensor, this rank's shard)."""
        m = self._bind()
        cfg = m._cfg
        L = _lib.lib()
        B = x.shape[0]
        if y.shape != (B, cfg.num_outputs) or y.dtype != torch.float32 or not y.is_contiguous():
            raise ValueError("y must be a contiguous float32 [B, num_outputs] tensor")
        pred, act = m._forward_impl(x, keep=False)
        self.pred = pred
        dpred = torch.empty_like(pred)
        _lib.check(L.sq_mse_fwd_bwd(_lib.ptr(pred), _lib.ptr(y), B, cfg.num_outputs, _lib.ptr(self.loss), _lib.ptr(dpred),
                                    _lib.ptr(self.mse_scratch), _lib.stream_ptr()))
        if self.world > 1 and self.overlap:
            # data parallel: a stage's gradient slice is all-reduced (NCCL) as soon as the stage is enqueued, and its AdamW
            # update runs on a side stream right after that all-reduce, while the backward pass continues below
            self.step_count += 1
            main = torch.cuda.current_stream()
            if self._opt_stream is None:
                self._opt_stream = torch.cuda.Stream()
            if self._mm is not None:
                # own all-reduce kernel on a communication stream; the GEMMs of the backward pass leave it `comm_ctas` SMs
                if self._comm_stream is None:
                    self._comm_stream = torch.cuda.Stream()
                L.sq_set_sm_budget(torch.cuda.get_device_properties(x.device).multi_processor_count - self.comm_ctas)
                try:
                    for s in range(cfg.depth, -1, -1):
                        m._backward_impl(act, dpred if s == cfg.depth else None, B, False, gbuf=self.g, stage_hi=s, stage_lo=s)
                        ev = torch.cuda.Event()
                        ev.record(main)
                        with torch.cuda.stream(self._comm_stream):
                            self._comm_stream.wait_event(ev)
                            self._mm.allreduce(*self.stage_range[s])
                            done = torch.cuda.Event()
                            done.record(self._comm_stream)
                        with torch.cuda.stream(self._opt_stream):
                            self._opt_stream.wait_event(done)
                            self._adamw(m, *self.stage_range[s])
                finally:
                    L.sq_set_sm_budget(0)
                main.wait_stream(self._opt_stream)
                m._planes_are_fresh()
                return self.loss
            for s in range(cfg.depth, -1, -1):
                m._backward_impl(act, dpred if s == cfg.depth else None, B, False, gbuf=self.g, stage_hi=s, stage_lo=s)
                w = allreduce_stage(self.g, self.stage_range[s], self.pg)
                with torch.cuda.stream(self._opt_stream):
                    w.wait()
                    self._adamw(m, *self.stage_range[s])
            main.wait_stream(self._opt_stream)
            m._planes_are_fresh()
            return self.loss
        elif self.world > 1 or not self.overlap:
            m._backward_impl(act, dpred, B, False, gbuf=self.g)
            if self.world > 1:
                torch.distributed.all_reduce(self.g, group=self.pg)
        else:
            # single GPU: the AdamW update of a stage (HBM-bound) runs on a side stream while the backward pass of the
            # stages below it (tensor-bound) continues; a stage's weights are not read again after its own backward
            self.step_count += 1
            main = torch.cuda.current_stream()
            if self._opt_stream is None:
                self._opt_stream = torch.cuda.Stream()
            for s in range(cfg.depth, -1, -1):
                m._backward_impl(act, dpred if s == cfg.depth else None, B, False, gbuf=self.g, stage_hi=s, stage_lo=s)
                ev = torch.cuda.Event()
                ev.record(main)
                with torch.cuda.stream(self._opt_stream):
                    self._opt_stream.wait_event(ev)
                    self._adamw(m, *self.stage_range[s])
            main.wait_stream(self._opt_stream)
            m._planes_are_fresh()
            return self.loss
        self.step_count += 1
        self._adamw(m, 0, m._total)
        m._planes_are_fresh()
        return self.loss

    def _adamw(self, m, b, e):
        off = 4 * b
        _lib.check(_lib.lib().sq_adamw_flat(C.c_void_p(m._flat.data_ptr() + off), C.c_void_p(self.g.data_ptr() + off),
                                            C.c_void_p(self.m.data_ptr() + off), C.c_void_p(self.v.data_ptr() + off),
                                            C.c_void_p(m._w_hi.data_ptr() + off // 2), C.c_void_p(m._w_lo.data_ptr() + off // 2), e - b, self.lr,
                                            self.betas[0], self.betas[1], self.eps, self.wd, self.step_count, 1.0 / self.world,
                                            _lib.stream_ptr()))


class HostBatchFeeder:
    """Double-buffered host -> device feed for the training loop (the reference's DataLoader uses pin_memory=True and moves each
    batch with `.to(model.device)`, src/main.py:120-135, src/vit.py:160-161): the H2D copy of batch i+1 runs on a copy stream
    while step i computes.  Iterate over it to get (x, y) device tensors that are safe to use on the current stream."""

    def __init__(self, batches, device, slots=2):
        self.batches = batches                      # sequence of (x_host, y_host) float32 CPU tensors (ideally pinned)
        self.device = torch.device(device)
        self.slots = slots
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.bufs = [None] * slots
        self.ready = [torch.cuda.Event() for _ in range(slots)]
        self.free = [torch.cuda.Event() for _ in range(slots)]
        self.h2d_bytes = 0

    def _issue(self, i):
        slot = i % self.slots
        xh, yh = self.batches[i]
        with torch.cuda.stream(self.copy_stream):
            if i >= self.slots:
                self.copy_stream.wait_event(self.free[slot])            # the step that used this slot has been enqueued and finished
            if self.bufs[slot] is None or self.bufs[slot][0].shape != xh.shape or self.bufs[slot][1].shape != yh.shape:
                self.bufs[slot] = (torch.empty(xh.shape, dtype=torch.float32, device=self.device),
                                   torch.empty(yh.shape, dtype=torch.float32, device=self.device))
            xd, yd = self.bufs[slot]
            xd.copy_(xh, non_blocking=True)
            yd.copy_(yh, non_blocking=True)
            self.ready[slot].record(self.copy_stream)
        self.h2d_bytes += xh.numel() * 4 + yh.numel() * 4

    def __iter__(self):
        n = len(self.batches)
        main = torch.cuda.current_stream(self.device)
        for i in range(min(self.slots - 1, n)):
            self._issue(i)
        for i in range(n):
            if i + self.slots - 1 < n:
                self._issue(i + self.slots - 1)
            slot = i % self.slots
            main.wait_event(self.ready[slot])
            yield self.bufs[slot]
            self.free[slot].record(main)
