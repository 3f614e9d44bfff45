"""Python-side helper around `sq_gemm_bf16` / `sq_split_bf16` (used by the tests and the UNI path)."""
import ctypes as C

import torch

from . import _lib

ACT = {"none": 0, "relu": 1, "gelu": 2, "ln64_gelu": 3, "mul_dgelu": 4}


def split_planes(x, ld_out=None, want_lo=True):
    """fp32 [rows, cols] -> (hi, lo) bf16 planes with row stride ld_out (default: cols rounded up to 8)."""
    assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    rows, cols = x.shape
    if ld_out is None:
        ld_out = (cols + 7) // 8 * 8
    hi = torch.zeros(rows, ld_out, dtype=torch.bfloat16, device=x.device)
    lo = torch.zeros(rows, ld_out, dtype=torch.bfloat16, device=x.device) if want_lo else None
    _lib.check(_lib.lib().sq_split_bf16(_lib.ptr(x), _lib.ptr(hi), _lib.ptr(lo), rows, cols, x.stride(0), ld_out,
                                        _lib.stream_ptr()))
    return hi, lo


def gemm(M, N, K, a_hi, b_hi, a_lo=None, b_lo=None, a_mn=False, b_mn=False, lda=None, ldb=None, nterms=1, split_k=0,
         block_n=0, a_koff_per_ntile=0, conv=None, out_f32=None, out_hi=None, out_lo=None, bias=None, rowbias=None,
         rowbias_div=1, res_f32=None, res_bf=None, save_pre=None, aux=None, ln_gamma=None, ln_beta=None, act="none",
         alpha=1.0, workspace=None, b_koff_per_ntile=0, b_nadj_per_ntile=0, b_map_mn=0, b_map_k=0, diag64=0):
    d = _lib.GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.a_hi, d.a_lo, d.a_mn_major = a_hi.data_ptr(), (a_lo.data_ptr() if a_lo is not None else None), int(a_mn)
    d.b_hi, d.b_lo, d.b_mn_major = b_hi.data_ptr(), (b_lo.data_ptr() if b_lo is not None else None), int(b_mn)
    d.lda = lda if lda is not None else a_hi.stride(0)
    d.ldb = ldb if ldb is not None else b_hi.stride(0)
    d.nterms, d.split_k, d.block_n, d.a_koff_per_ntile = nterms, split_k, block_n, a_koff_per_ntile
    if split_k > 1 and workspace is None:
        workspace = torch.empty(split_k * M * N, dtype=torch.float32, device=a_hi.device)
    if workspace is not None:
        d.workspace, d.workspace_bytes = workspace.data_ptr(), workspace.numel() * workspace.element_size()
    if conv is not None:
        d.conv_enabled = 1
        (d.conv_batch, d.conv_H, d.conv_W, d.conv_C, d.conv_Ho, d.conv_Wo, d.conv_R, d.conv_S, d.conv_stride,
         d.conv_pad) = conv
    for name, t in (("out_f32", out_f32), ("out_hi", out_hi), ("out_lo", out_lo), ("bias", bias), ("rowbias", rowbias),
                    ("res_f32", res_f32), ("res_bf", res_bf), ("save_pre", save_pre), ("aux", aux),
                    ("ln_gamma", ln_gamma), ("ln_beta", ln_beta)):
        if t is not None:
            setattr(d, name, t.data_ptr())
    if out_f32 is not None:
        d.ld_f32 = out_f32.stride(0)
    if out_hi is not None:
        d.ld_bf = out_hi.stride(0)
    if rowbias is not None:
        d.rowbias_div, d.ld_rowbias = rowbias_div, rowbias.stride(0)
    if res_f32 is not None:
        d.ld_res = res_f32.stride(0)
    if res_bf is not None:
        d.ld_res = res_bf.stride(0)
    if save_pre is not None:
        d.ld_pre = save_pre.stride(0)
    if aux is not None:
        d.ld_aux = aux.stride(0)
    d.act, d.alpha = ACT[act], alpha
    d.b_koff_per_ntile, d.b_nadj_per_ntile, d.b_map_mn, d.b_map_k, d.diag64 = (b_koff_per_ntile, b_nadj_per_ntile,
                                                                               b_map_mn, b_map_k, diag64)
    _lib.check(_lib.lib().sq_gemm_bf16(C.byref(d), _lib.stream_ptr()))
    return workspace


def conv_bf16(x, w, shift, residual=None, relu=True, stride=1, pad=0, block_n=0, cta_group=0, out=None):
    """x: bf16 NHWC [B,H,W,Cin]; w: bf16 [Cout,R,S,Cin]; shift: fp32 [Cout]; residual: bf16 [B,Ho,Wo,Cout] -> bf16 [B,Ho,Wo,Cout]
    through `sq_conv_bf16` (the CTA-pair tcgen05 kernel with the TMA epilogue)."""
    B, H, W, Cin = x.shape
    Cout, R, S, _ = w.shape
    Ho, Wo = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
    if out is None:
        out = torch.empty(B, Ho, Wo, Cout, dtype=torch.bfloat16, device=x.device)
    d = _lib.ConvDesc()
    d.batch, d.H, d.W, d.Cin, d.Cout, d.R, d.S, d.stride, d.pad = B, H, W, Cin, Cout, R, S, stride, pad
    d.inp, d.weight, d.shift, d.out = x.data_ptr(), w.data_ptr(), shift.data_ptr(), out.data_ptr()
    d.residual = residual.data_ptr() if residual is not None else None
    d.relu, d.block_n, d.cta_group = int(relu), block_n, cta_group
    _lib.check(_lib.lib().sq_conv_bf16(C.byref(d), _lib.stream_ptr()))
    return out


def bneck_l1_bf16(x, w2, shift2, w3, shift3, residual, out=None):
    """Fused tail of a layer-1 bottleneck through `sq_bneck_l1_bf16`: x bf16 NHWC [B,H,W,64], w2 bf16 [64,3,3,64], w3 bf16 [256,64] (or
    [256,1,1,64]), shifts fp32, residual bf16 [B,H,W,256] -> relu(conv1x1(relu(conv3x3(x) + shift2)) + shift3 + residual)."""
    B, H, W, _ = x.shape
    if out is None:
        out = torch.empty(B, H, W, 256, dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.lib().sq_bneck_l1_bf16(C.c_void_p(x.data_ptr()), C.c_void_p(w2.data_ptr()), C.c_void_p(shift2.data_ptr()), C.c_void_p(w3.data_ptr()),
                                           C.c_void_p(shift3.data_ptr()), C.c_void_p(residual.data_ptr()), C.c_void_p(out.data_ptr()), B, H, W,
                                           _lib.stream_ptr()))
    return out


def bneck_l1_ds_bf16(x1, w2, shift2, w3, shift3, x, wds, shiftds, out=None):
    """First-block variant through `sq_bneck_l1_ds_bf16`: x1 = conv1's output [B,H,W,64], x = the block input [B,H,W,64], wds bf16 [256,64]
    -> relu(conv1x1(relu(conv3x3(x1) + shift2)) + shift3 + conv1x1(x, wds) + shiftds), the downsample computed inside the kernel."""
    B, H, W, _ = x1.shape
    if out is None:
        out = torch.empty(B, H, W, 256, dtype=torch.bfloat16, device=x1.device)
    p = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(_lib.lib().sq_bneck_l1_ds_bf16(p(x1), p(w2), p(shift2), p(w3), p(shift3), p(x), p(wds), p(shiftds), p(out), B, H, W, _lib.stream_ptr()))
    return out

