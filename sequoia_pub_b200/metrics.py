"""Device-side versions of the per-step metrics of the reference training loop (src/vit.py:167-168):
`sklearn.metrics.mean_absolute_error(labels, preds)` and `he2rna.compute_correlations(labels, preds)`
(src/he2rna.py:140-149), plus evaluate()'s `smape` (src/vit.py:32-33,269).  The reference copies pred / labels to the host and loops over ~20k genes in numpy every step;
here one kernel produces both numbers without leaving the device.  No CPU fallback."""
import numpy as np
import torch

from . import _lib


def step_metrics(labels, preds):
    """labels, preds: float32 [B, G] CUDA tensors -> float32[4] device tensor (MAE, mean per-gene Pearson r, #genes used, SMAPE)."""
    if isinstance(labels, np.ndarray):
        labels = torch.from_numpy(np.ascontiguousarray(labels, dtype=np.float32)).cuda()
    if isinstance(preds, np.ndarray):
        preds = torch.from_numpy(np.ascontiguousarray(preds, dtype=np.float32)).cuda()
    if labels.shape != preds.shape or labels.dim() != 2 or labels.dtype != torch.float32 or preds.dtype != torch.float32:
        raise ValueError("labels and preds must be float32 [B, G] tensors of the same shape")
    if not labels.is_cuda or not preds.is_cuda:
        raise RuntimeError("sequoia_b200 metrics run on the GPU only (no CPU fallback)")
    _lib.require_device()
    labels, preds = labels.contiguous(), preds.contiguous()
    B, G = labels.shape
    L = _lib.lib()
    scratch = torch.empty(L.sq_step_metrics_scratch_bytes(G), dtype=torch.uint8, device=labels.device)
    out = torch.empty(4, dtype=torch.float32, device=labels.device)
    with _lib.on_device(labels):
        _lib.check(L.sq_step_metrics(_lib.ptr(labels), _lib.ptr(preds), B, G, _lib.ptr(out), _lib.ptr(scratch), scratch.numel(),
                                     _lib.stream_ptr(labels)))
    return out


def compute_correlations(labels, preds):
    """Same signature / return value as src/he2rna.py:140-149 (a Python float); accepts numpy arrays or CUDA tensors."""
    return float(step_metrics(labels, preds)[1].item())


def mean_absolute_error(labels, preds):
    return float(step_metrics(labels, preds)[0].item())


def smape(labels, preds):
    """`smape(A, F)` of src/vit.py:32-33 as evaluate() calls it (:269): 100 / len(A) * sum(2 |F - A| / (|A| + |F|)) over the whole
    [B, G] array, len(A) = B."""
    return float(step_metrics(labels, preds)[3].item())
