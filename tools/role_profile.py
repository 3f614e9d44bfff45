"""Per-warp-role cycle breakdown of the GEMM kernel on a few representative shapes (development aid)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequoia_pub_b200 import _gemm as gm, _lib

L = _lib.lib()
prof = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")

def report(name):
    torch.cuda.synchronize()
    p = prof.view(148, 16).double()
    act = p[:, 4] > 0
    m = p[act].mean(0)
    print(f"{name}: producer wait_empty {m[0]:.0f} / {m[1]:.0f} | mma wait_full {m[2]:.0f} wait_tmem_empty {m[3]:.0f} / {m[4]:.0f} | "
          f"epi(w4) wait_tmem_full {m[5]:.0f} / {m[13]:.0f}  (cycles, mean over {int(act.sum())} CTAs)")
    prof.zero_()

def run(fn, name):
    for _ in range(2): fn()
    L.sq_gemm_profile(_lib.ptr(prof)); prof.zero_(); fn(); report(name); L.sq_gemm_profile(None)

g = torch.Generator(device="cuda").manual_seed(0)
# conv3 of layer1: M=262144 (64 imgs x 64x64), N=256, K=64, bf16 residual, relu, bf16 out
x = torch.randn(64, 64, 64, 64, device="cuda", generator=g).to(torch.bfloat16)
w = (torch.randn(256, 64, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
res = torch.randn(262144, 256, device="cuda", generator=g).to(torch.bfloat16)
out = torch.empty(262144, 256, device="cuda", dtype=torch.bfloat16)
shift = torch.randn(256, device="cuda")
conv = (64, 64, 64, 64, 64, 64, 1, 1, 1, 0)
run(lambda: gm.gemm(262144, 256, 64, x.view(-1, 64), w, conv=conv, out_hi=out, bias=shift, res_bf=res, act="relu"), "resnet layer1 conv3 (N=256,K=64,res)")
run(lambda: gm.gemm(262144, 256, 64, x.view(-1, 64), w, conv=conv, out_hi=out, bias=shift, act="relu"), "  same, no residual")
run(lambda: gm.gemm(262144, 256, 64, x.view(-1, 64), w, conv=conv, out_hi=out, bias=shift, act="relu", block_n=128), "  same, no residual, BN=128")
w1 = (torch.randn(64, 64, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
out1 = torch.empty(262144, 64, device="cuda", dtype=torch.bfloat16)
run(lambda: gm.gemm(262144, 64, 64, x.view(-1, 64), w1, conv=conv, out_hi=out1, bias=shift[:64], act="relu"), "resnet layer1 conv1 (N=64,K=64)")
# ViS FF GEMM
A = torch.randn(3200, 2048, device="cuda", generator=g); B = torch.randn(2048, 2048, device="cuda", generator=g)
a_hi, a_lo = gm.split_planes(A); b_hi, b_lo = gm.split_planes(B)
o32 = torch.empty(3200, 2048, device="cuda"); ohi = torch.empty(3200, 2048, device="cuda", dtype=torch.bfloat16); olo = torch.empty_like(ohi); pre = torch.empty_like(o32)
bias = torch.randn(2048, device="cuda")
for bn in (128, 256):
    run(lambda: gm.gemm(3200, 2048, 2048, a_hi, b_hi, a_lo, b_lo, nterms=3, out_hi=ohi, out_lo=olo, save_pre=pre, bias=bias, act="gelu", block_n=bn), f"vis W1 (3200x2048x2048 x3, gelu+planes+pre) BN={bn}")
    run(lambda: gm.gemm(3200, 2048, 2048, a_hi, b_hi, a_lo, b_lo, nterms=3, out_f32=o32, block_n=bn), f"vis plain f32 out BN={bn}")
