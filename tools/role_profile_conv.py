"""Per-warp-role cycle counters of convgemm_kernel (csrc/convgemm.cuh) on the ResNet shapes, plus CUDA-event timings.
   python tools/role_profile_conv.py [cta_group]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequoia_pub_b200 import _gemm as gm, _lib

L = _lib.lib()
cg = int(sys.argv[1]) if len(sys.argv) > 1 else 0
prof = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
g = torch.Generator(device="cuda").manual_seed(0)

def bench(fn, name, flops, bytes_):
    for _ in range(2): fn()
    ts = []
    for _ in range(5):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e) * 1e3)
    L.sq_gemm_profile(_lib.ptr(prof)); prof.zero_(); flush.zero_(); fn(); torch.cuda.synchronize(); L.sq_gemm_profile(None)
    p = prof.view(148, 16).double(); act = p[:, 1] > 0; m = p[act].mean(0)
    t = sorted(ts)[len(ts) // 2]
    print(f"{name:34s} {t:7.1f} us  {flops / t / 1e6:6.0f} TF/s {bytes_ / t / 1e3:6.0f} GB/s | prod wait {m[0]:7.0f}/{m[1]:7.0f} | mma wait_full {m[2]:7.0f} wait_acc_free {m[3]:7.0f} /{m[4]:7.0f} | "
          f"epi0 wait_acc {m[5]:7.0f} wait_res {m[6]:7.0f} wait_store {m[7]:6.0f} /{m[8]:7.0f}", flush=True)

def conv_case(name, B, H, Cin, Cout, k, stride, res, bn=0):
    pad = k // 2
    x = torch.randn(B, H, H, Cin, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(Cout, k, k, Cin, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    sh = torch.randn(Cout, device="cuda", generator=g)
    Ho = (H + 2 * pad - k) // stride + 1
    r = torch.randn(B, Ho, Ho, Cout, device="cuda", generator=g).to(torch.bfloat16) if res else None
    out = torch.empty(B, Ho, Ho, Cout, device="cuda", dtype=torch.bfloat16)
    M = B * Ho * Ho
    flops = 2.0 * M * Cout * Cin * k * k
    bytes_ = 2.0 * (x.numel() + out.numel() * (2 if res else 1) + w.numel())
    bench(lambda: gm.conv_bf16(x, w, sh, r, True, stride, pad, bn, cg, out), name, flops, bytes_)

conv_case("L1 c1 64->64 1x1", 64, 64, 64, 64, 1, 1, False)
conv_case("L1 c2 64->64 3x3", 64, 64, 64, 64, 3, 1, False)
conv_case("L1 ds 64->256 1x1", 64, 64, 64, 256, 1, 1, False)
conv_case("L1 c3 64->256 1x1 +res", 64, 64, 64, 256, 1, 1, True)
conv_case("L1 c3 same BN=128", 64, 64, 64, 256, 1, 1, True, 128)
conv_case("L1 c1' 256->64 1x1", 64, 64, 256, 64, 1, 1, False)
conv_case("L2 c1 256->128 1x1", 64, 64, 256, 128, 1, 1, False)
conv_case("L2 c2 128->128 3x3 s2", 64, 64, 128, 128, 3, 2, False)
conv_case("L2 c2 128->128 3x3", 64, 32, 128, 128, 3, 1, False)
conv_case("L2 c3 128->512 +res", 64, 32, 128, 512, 1, 1, True)
conv_case("L2 c1' 512->128", 64, 32, 512, 128, 1, 1, False)
conv_case("L3 c2 256->256 3x3", 64, 16, 256, 256, 3, 1, False)
conv_case("L3 c3 256->1024 +res", 64, 16, 256, 1024, 1, 1, True)
conv_case("L3 c1' 1024->256", 64, 16, 1024, 256, 1, 1, False)
conv_case("L4 c2 512->512 3x3", 64, 8, 512, 512, 3, 1, False)
conv_case("L4 c3 512->2048 +res", 64, 8, 512, 2048, 1, 1, True)
conv_case("L4 c1' 2048->512", 64, 8, 2048, 512, 1, 1, False)
