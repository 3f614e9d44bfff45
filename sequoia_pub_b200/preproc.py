"""GPU image preprocessing of the extraction scripts.

`resize_tiles` = `transforms.Resize(size)` applied to `Image.fromarray(tile).convert("RGB")` for every tile
(pre_processing/compute_features_hdf5.py:53-56,125-126): Pillow's antialiased bilinear resample, reproduced bit for bit by
`sq_resize_bilinear_u8` (csrc/preproc.cu).  The rest of that transform (ToTensor's /255 and Normalize) is fused into the first
kernel of the extractors.  No CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

_tables = {}


def resize_size(h, w, size=224):
    """Output (h', w') of torchvision's `Resize(int)`: smaller edge -> size, aspect ratio kept (truncated)."""
    if h <= w:
        return size, int(size * w / h)
    return int(size * h / w), size


def coeff_tables(in_size, out_size):
    """Host int32 arrays (bounds [out, 2], weights [out, ksize]) from the library's restatement of Resample.c precompute_coeffs."""
    L = _lib.lib()
    ks = L.sq_resize_ksize(in_size, out_size)
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ks), np.int32)
    _lib.check(L.sq_resize_coeffs(in_size, out_size, bounds.ctypes.data_as(C.POINTER(C.c_int)), kk.ctypes.data_as(C.POINTER(C.c_int))))
    return bounds, kk


def _device_tables(in_size, out_size, device):
    key = (in_size, out_size, str(device))
    if key not in _tables:
        b, k = coeff_tables(in_size, out_size)
        _tables[key] = (torch.from_numpy(b).to(device), torch.from_numpy(k).to(device), k.shape[1])
    return _tables[key]


def resize_scratch_bytes(n, h, w, size=224):
    oh, ow = resize_size(h, w, size)
    return n * h * ow * 3 if (oh != h and ow != w) else 0


@torch.no_grad()
def resize_tiles(tiles, size=224, out=None, tmp=None):
    """tiles: CUDA uint8 [n, H, W, 3] -> CUDA uint8 [n, h', w', 3] (enqueued on the current stream of the tiles' device)."""
    if tiles.dim() != 4 or tiles.shape[3] != 3 or tiles.dtype != torch.uint8 or not tiles.is_cuda:
        raise ValueError("resize_tiles expects a CUDA uint8 [n,H,W,3] tensor")
    tiles = tiles.contiguous()
    n, h, w = tiles.shape[0], tiles.shape[1], tiles.shape[2]
    oh, ow = resize_size(h, w, size)
    if out is None:
        out = torch.empty(n, oh, ow, 3, dtype=torch.uint8, device=tiles.device)
    elif tuple(out.shape) != (n, oh, ow, 3) or out.dtype != torch.uint8 or not out.is_contiguous():
        raise ValueError("out must be a contiguous uint8 [n,h',w',3] tensor")
    xb = xk = yb = yk = None
    xks = yks = 0
    if ow != w:
        xb, xk, xks = _device_tables(w, ow, tiles.device)
    if oh != h:
        yb, yk, yks = _device_tables(h, oh, tiles.device)
    need = resize_scratch_bytes(n, h, w, size)
    if need and (tmp is None or tmp.numel() < need):
        tmp = torch.empty(need, dtype=torch.uint8, device=tiles.device)
    with _lib.on_device(tiles):
        _lib.check(_lib.lib().sq_resize_bilinear_u8(_lib.ptr(tiles), n, h, w, _lib.ptr(out), oh, ow, _lib.ptr(xb), _lib.ptr(xk), xks,
                                                    _lib.ptr(yb), _lib.ptr(yk), yks, _lib.ptr(tmp), _lib.stream_ptr(tiles)))
    return out
