"""Per-warp-role cycle breakdown for the ResNet layer-1/stem shapes (development aid)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequoia_pub_b200 import _gemm as gm, _lib
L = _lib.lib()
prof = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
def run(fn, name):
    for _ in range(2): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); fn(); e.record(); torch.cuda.synchronize(); us = s.elapsed_time(e) * 1e3
    L.sq_gemm_profile(_lib.ptr(prof)); prof.zero_(); fn(); torch.cuda.synchronize(); L.sq_gemm_profile(None)
    p = prof.view(148, 16).double(); m = p[p[:, 4] > 0].mean(0)
    print(f"{name}: {us:.1f} us | producer wait_empty {m[0]:.0f}/{m[1]:.0f} | mma wait_full {m[2]:.0f} wait_tmem_empty {m[3]:.0f} / {m[4]:.0f} | epi(w4) wait {m[5]:.0f} / {m[13]:.0f}")
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(64, 64, 64, 64, device="cuda", generator=g).to(torch.bfloat16)
w3 = (torch.randn(64, 576, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
out1 = torch.empty(262144, 64, device="cuda", dtype=torch.bfloat16); shift = torch.randn(256, device="cuda")
run(lambda: gm.gemm(262144, 64, 576, x.view(-1, 64), w3, conv=(64, 64, 64, 64, 64, 64, 3, 3, 1, 1), out_hi=out1, bias=shift[:64], act="relu"), "layer1 conv2 3x3 64->64")
w1 = (torch.randn(64, 64, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
run(lambda: gm.gemm(262144, 64, 64, x.view(-1, 64), w1, conv=(64, 64, 64, 64, 64, 64, 1, 1, 1, 0), out_hi=out1, bias=shift[:64], act="relu"), "layer1 conv1 1x1 64->64")
col = torch.randn(1048576, 192, device="cuda", generator=g).to(torch.bfloat16); ws = (torch.randn(64, 192, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
outs = torch.empty(1048576, 64, device="cuda", dtype=torch.bfloat16)
run(lambda: gm.gemm(1048576, 64, 192, col, ws, out_hi=outs, bias=shift[:64], act="relu"), "stem GEMM 1M x 64 x 192")
x2 = torch.randn(64, 32, 32, 128, device="cuda", generator=g).to(torch.bfloat16)
w32 = (torch.randn(128, 1152, device="cuda", generator=g) * 0.05).to(torch.bfloat16); out2 = torch.empty(65536, 128, device="cuda", dtype=torch.bfloat16)
run(lambda: gm.gemm(65536, 128, 1152, x2.view(-1, 128), w32, conv=(64, 32, 32, 128, 32, 32, 3, 3, 1, 1), out_hi=out2, bias=shift[:128], act="relu"), "layer2 conv2 3x3 128->128")
run(lambda: gm.gemm(262144, 64, 64, x.view(-1, 64), w1, out_hi=out1, bias=shift[:64], act="relu"), "layer1 conv1 as PLAIN 2-D GEMM")
w256 = (torch.randn(256, 64, device="cuda", generator=g) * 0.05).to(torch.bfloat16); out256 = torch.empty(262144, 256, device="cuda", dtype=torch.bfloat16)
res = torch.randn(262144, 256, device="cuda", generator=g).to(torch.bfloat16)
run(lambda: gm.gemm(262144, 256, 64, x.view(-1, 64), w256, conv=(64, 64, 64, 64, 64, 64, 1, 1, 1, 0), out_hi=out256, bias=shift, res_bf=res, act="relu"), "layer1 conv3 (conv mode)")
run(lambda: gm.gemm(262144, 256, 64, x.view(-1, 64), w256, out_hi=out256, bias=shift, res_bf=res, act="relu"), "layer1 conv3 as PLAIN 2-D GEMM")
