"""ctypes binding of libsequoia_b200.so (the C ABI declared in include/sequoia_b200.h).

There is no fallback: if the shared library is missing the import of any product module fails,
and every call that returns non-zero raises RuntimeError(sq_last_error()).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsequoia_b200.so")

c_void_p, c_int, c_ll, c_float, c_size_t = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t


class GemmDesc(C.Structure):
    """Mirror of `sq_gemm_desc`."""
    _fields_ = [
        ("M", c_int), ("N", c_int), ("K", c_int),
        ("a_hi", c_void_p), ("a_lo", c_void_p), ("a_mn_major", c_int), ("lda", c_ll),
        ("b_hi", c_void_p), ("b_lo", c_void_p), ("b_mn_major", c_int), ("ldb", c_ll),
        ("nterms", c_int), ("split_k", c_int), ("block_n", c_int), ("a_koff_per_ntile", c_int),
        ("workspace", c_void_p), ("workspace_bytes", c_size_t),
        ("conv_enabled", c_int), ("conv_batch", c_int), ("conv_H", c_int), ("conv_W", c_int), ("conv_C", c_int),
        ("conv_Ho", c_int), ("conv_Wo", c_int), ("conv_R", c_int), ("conv_S", c_int), ("conv_stride", c_int),
        ("conv_pad", c_int),
        ("out_f32", c_void_p), ("ld_f32", c_ll),
        ("out_hi", c_void_p), ("out_lo", c_void_p), ("ld_bf", c_ll),
        ("bias", c_void_p),
        ("rowbias", c_void_p), ("rowbias_div", c_int), ("ld_rowbias", c_ll),
        ("res_f32", c_void_p), ("res_bf", c_void_p), ("ld_res", c_ll),
        ("save_pre", c_void_p), ("ld_pre", c_ll),
        ("aux", c_void_p), ("ld_aux", c_ll),
        ("ln_gamma", c_void_p), ("ln_beta", c_void_p),
        ("act", c_int), ("alpha", c_float),
        ("b_koff_per_ntile", c_int), ("b_nadj_per_ntile", c_int), ("b_map_mn", c_int), ("b_map_k", c_int),
        ("diag64", c_int),
    ]


class ConvDesc(C.Structure):
    """Mirror of `sq_conv_desc`."""
    _fields_ = [("batch", c_int), ("H", c_int), ("W", c_int), ("Cin", c_int), ("Cout", c_int), ("R", c_int), ("S", c_int),
                ("stride", c_int), ("pad", c_int),
                ("inp", c_void_p), ("weight", c_void_p), ("shift", c_void_p), ("residual", c_void_p), ("out", c_void_p),
                ("relu", c_int), ("block_n", c_int), ("cta_group", c_int)]


class VisConfig(C.Structure):
    """Mirror of `sq_vis_config`."""
    _fields_ = [("input_dim", c_int), ("depth", c_int), ("nheads", c_int), ("num_clusters", c_int), ("num_outputs", c_int)]


class VitConfig(C.Structure):
    """Mirror of `sq_vit_config` (the first field is `dim` in C; named like VisConfig's so the shared host code reads both)."""
    _fields_ = [("input_dim", c_int), ("depth", c_int), ("nheads", c_int), ("num_clusters", c_int), ("num_outputs", c_int),
                ("mlp_dim", c_int)]


# name -> (restype, argtypes); kept in one table so tests can check the export list against the header
SIGNATURES = {
    "sq_version": (c_int, []),
    "sq_last_error": (C.c_char_p, []),
    "sq_device_ok": (c_int, []),
    "sq_gemm_timing_enable": (c_int, [c_int]),
    "sq_gemm_timing_read": (c_int, [C.POINTER(C.c_double), C.POINTER(c_ll), C.POINTER(C.c_double)]),
    "sq_gemm_profile": (c_int, [c_void_p]),
    "sq_side_stream_enable": (c_int, [c_int]),
    "sq_set_sm_budget": (c_int, [c_int]),
    "sq_multimem_flag_bytes": (c_size_t, [c_int]),
    "sq_multimem_allreduce_f32": (c_int, [c_void_p, c_ll, c_ll, C.POINTER(c_void_p), c_int, c_int, C.c_uint, c_int, c_void_p]),
    "sq_split_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_ll, c_int, c_ll, c_ll, c_void_p]),
    "sq_gemm_bf16": (c_int, [C.POINTER(GemmDesc), c_void_p]),
    "sq_conv_bf16": (c_int, [C.POINTER(ConvDesc), c_void_p]),
    "sq_bneck_l1_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "sq_bneck_l1_ds_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "sq_resnet50_num_convs": (c_int, []),
    "sq_resnet50_conv_info": (c_int, [c_int] + [C.POINTER(c_int)] * 5),
    "sq_resnet50_packed_weight_elems": (c_ll, []),
    "sq_resnet50_shift_elems": (c_ll, []),
    "sq_resnet50_prepack": (c_int, [C.POINTER(c_void_p), c_void_p, c_void_p, c_float, c_void_p]),
    "sq_resnet50_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "sq_resnet50_extract": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_size_t, c_void_p]),
    "sq_resnet50_prepack_planes": (c_int, [C.POINTER(c_void_p), c_void_p, c_void_p, c_void_p, c_float, c_void_p]),
    "sq_resnet50_hp_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "sq_resnet50_extract_hp": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_size_t, c_void_p]),
    "sq_vis_param_table_len": (c_int, [C.POINTER(VisConfig)]),
    "sq_vis_param_layout": (c_int, [C.POINTER(VisConfig), C.POINTER(c_ll), c_int, C.POINTER(c_ll)]),
    "sq_vis_act_bytes": (c_size_t, [C.POINTER(VisConfig), c_int]),
    "sq_vis_bwd_bytes": (c_size_t, [C.POINTER(VisConfig), c_int]),
    "sq_vis_forward": (c_int, [C.POINTER(VisConfig), c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                               c_size_t, c_void_p]),
    "sq_vis_backward": (c_int, [C.POINTER(VisConfig), c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_size_t,
                                c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_int, c_void_p]),
    "sq_vit_param_table_len": (c_int, [C.POINTER(VitConfig)]),
    "sq_vit_param_layout": (c_int, [C.POINTER(VitConfig), C.POINTER(c_ll), c_int, C.POINTER(c_ll)]),
    "sq_vit_act_bytes": (c_size_t, [C.POINTER(VitConfig), c_int]),
    "sq_vit_bwd_bytes": (c_size_t, [C.POINTER(VitConfig), c_int]),
    "sq_vit_forward": (c_int, [C.POINTER(VitConfig), c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                               c_size_t, c_void_p]),
    "sq_vit_backward": (c_int, [C.POINTER(VitConfig), c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_size_t,
                                c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_int, c_void_p]),
    "sq_mse_fwd_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sq_adamw_flat": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_float, c_float, c_float,
                              c_float, c_float, c_int, c_float, c_void_p]),
    "sq_step_metrics_scratch_bytes": (c_size_t, [c_int]),
    "sq_step_metrics": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "sq_vitl16_num_tensors": (c_int, [c_int]),
    "sq_vitl16_packed_weight_elems": (c_ll, [c_int]),
    "sq_vitl16_packed_vec_elems": (c_ll, [c_int]),
    "sq_vitl16_prepack": (c_int, [C.POINTER(c_void_p), c_int, c_void_p, c_void_p, c_void_p]),
    "sq_vitl16_workspace_bytes": (c_size_t, [c_int]),
    "sq_vitl16_extract": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "sq_resize_ksize": (c_int, [c_int, c_int]),
    "sq_resize_coeffs": (c_int, [c_int, c_int, C.POINTER(c_int), C.POINTER(c_int)]),
    "sq_resize_bilinear_u8": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                      c_int, c_void_p, c_void_p]),
    "sq_kmeans_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "sq_kmeans_fit": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_size_t, c_void_p]),
}

_lib = None


def lib():
    """Loads the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(sequoia_b200 has no CPU / PyTorch fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError("sequoia_b200: " + lib().sq_last_error().decode("utf-8", "replace"))


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else c_void_p(t.data_ptr())


def stream_ptr(t=None):
    """Current CUDA stream of the device that holds tensor `t` (or of the current device).  The library enqueues on the device
    that is current when it is called, so callers whose tensors may live on another device wrap the call in `on_device(t)`."""
    import torch
    dev = t.device if t is not None and getattr(t, "is_cuda", False) else None
    return c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def on_device(t):
    """Context manager that makes the device of tensor `t` current (models built with device='cuda:1' and no set_device)."""
    import torch
    return torch.cuda.device(t.device if getattr(t, "is_cuda", False) else torch.cuda.current_device())


def with_device_of(pick):
    """Decorator: runs the wrapped method with the CUDA device of `pick(*args, **kwargs)` (a tensor) current, so that every
    `stream_ptr()` / `require_device()` inside refers to the device that holds the operands (a model built with device='cuda:1' in a
    process whose current device is 0 must not enqueue on device 0's stream)."""
    import functools

    def deco(fn):
        @functools.wraps(fn)
        def wrapper(*args, **kwargs):
            t = pick(*args, **kwargs)
            if t is None or not getattr(t, "is_cuda", False):
                return fn(*args, **kwargs)
            with on_device(t):
                return fn(*args, **kwargs)
        return wrapper
    return deco


_device_checked = set()


def require_device():
    """Fails loudly unless the current CUDA device is an sm_100 part (checked once per device: the query is slow)."""
    import torch
    dev = torch.cuda.current_device()
    if dev not in _device_checked:
        check(lib().sq_device_ok())
        _device_checked.add(dev)
