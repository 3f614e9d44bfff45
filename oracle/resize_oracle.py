"""ORACLE (test infrastructure, not product code) — CPU restatement of the image resize in front of UNI.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

Reference call site: pre_processing/compute_features_hdf5.py:53-56,125-126
    Image.fromarray(tile).convert("RGB") -> transforms.Resize(224) -> ToTensor -> Normalize
`transforms.Resize(224)` on a PIL image is `Image.resize((w', h'), BILINEAR)` (torchvision.transforms.functional._pil), whose
arithmetic lives in the third-party dependency Pillow (pinned `pillow==10.3.0` in the reference's requirements.txt; 12.2.0 is what
this image has; libImaging/Resample.c is unchanged between them for this path).  Restated here from Resample.c:
  * precompute_coeffs: per output pixel a window [xmin, xmin + xmax) around centre (x + 0.5) * scale, triangle filter of
    support `scale` (antialiasing when shrinking), weights normalised in double;
  * normalize_coeffs_8bpc: weights to fixed point with 22 fractional bits, rounded half away from zero;
  * ImagingResampleHorizontal_8bpc then ImagingResampleVertical_8bpc: integer accumulation starting from 1 << 21, arithmetic
    shift by 22, clamp to [0, 255] — the horizontal pass is rounded to uint8 before the vertical pass.
Pinned against Pillow itself in tests/test_oracle_cpu.py (bit-exact on random tiles).
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def resize_size(h, w, size=224):
    """torchvision `Resize(int)`: the smaller edge becomes `size`, the other keeps the aspect ratio (truncated)."""
    if h <= w:
        return size, int(size * w / h)
    return int(size * h / w), size


def coeffs(in_size, out_size):
    """(bounds int32 [out, 2] = (first input index, count), weights int32 [out, ksize]) of Resample.c for the bilinear filter."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        xmin = max(xmin, 0)
        xmax = int(center + support + 0.5)
        xmax = min(xmax, in_size) - xmin
        w = []
        for x in range(xmax):
            t = abs((x + xmin - center + 0.5) * ss)
            w.append(1.0 - t if t < 1.0 else 0.0)
        ww = sum(w)                                   # accumulated left to right in double, like the C loop
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img, bounds, kk, axis):
    """One separable pass over `axis` of a uint8 [H, W, C] image."""
    img = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((bounds.shape[0],) + img.shape[1:], np.uint8)
    for i in range(bounds.shape[0]):
        x0, n = int(bounds[i, 0]), int(bounds[i, 1])
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(n):
            acc += img[x0 + x] * int(kk[i, x])
        out[i] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize(tile, size=224):
    """uint8 [H, W, 3] -> uint8 [h', w', 3] exactly as `transforms.Resize(size)(Image.fromarray(tile).convert("RGB"))`."""
    h, w = tile.shape[:2]
    oh, ow = resize_size(h, w, size)
    out = tile
    if ow != w:
        out = _pass(out, *coeffs(w, ow), axis=1)       # horizontal first (Resample.c ImagingResampleInner)
    if oh != h:
        out = _pass(out, *coeffs(h, oh), axis=0)
    return out
