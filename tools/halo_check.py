"""Halo-staged 3x3 convolutions (csrc/convgemm.cuh, HALO = 1) against the per-tap mode: bit-exact check + timing.
   python tools/halo_check.py      (each arm in its own interpreter: SQ_CONV_HALO is read once)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARM = r'''
import sys, torch
sys.path.insert(0, %r)
from sequoia_pub_b200 import _gemm as gm
g = torch.Generator(device="cuda").manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
outs = []
for (B, H, C, N) in ((64, 64, 64, 64), (64, 32, 128, 128), (64, 16, 256, 256), (3, 32, 128, 128), (5, 16, 64, 64)):
    x = torch.randn(B, H, H, C, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, 3, 3, C, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    sh = torch.randn(N, device="cuda", generator=g)
    o = gm.conv_bf16(x, w, sh, None, True, 1, 1)
    ts = []
    for _ in range(5):
        flush.zero_(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); gm.conv_bf16(x, w, sh, None, True, 1, 1, out=o); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e) * 1e3)
    print(f"  B={B} H={H} C={C} N={N}: {sorted(ts)[2]:.1f} us")
    outs.append(o.float().cpu())
torch.save(outs, sys.argv[1])
''' % ROOT
if __name__ == "__main__":
    import torch
    res = {}
    for mode in ("0", "1", "2"):
        path = f"/tmp/halo_{mode}.pt"
        r = subprocess.run([sys.executable, "-c", ARM, path], env=dict(os.environ, SQ_CONV_HALO=mode), capture_output=True, text=True, timeout=300)
        print(f"SQ_CONV_HALO={mode}: rc={r.returncode}\n{r.stdout.rstrip()} {r.stderr.strip()[-500:]}", flush=True)
        res[mode] = torch.load(path) if r.returncode == 0 else None
    for mode in ("1", "2"):
        if res["0"] is not None and res[mode] is not None:
            print("halo mode", mode, "vs per-tap:", [(bool(torch.equal(a, b)), float((a - b).abs().max())) for a, b in zip(res["0"], res[mode])])
