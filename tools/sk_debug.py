import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequoia_pub_b200 import _gemm as gm
torch.manual_seed(0)
ws = torch.zeros(160 * 256 * 128 + 4096, dtype=torch.float32, device="cuda")
M, N, K = 3200, 2048, 2048
A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda")
a_hi, a_lo = gm.split_planes(A); b_hi, b_lo = gm.split_planes(B.t().contiguous())
aux = torch.randn(M, N, device="cuda")
ad = aux.double().requires_grad_(True); torch.nn.functional.gelu(ad).sum().backward()
ref64 = (A.double() @ B.double().t()) * ad.grad
for name, kw in [("streamk bn256", dict(workspace=ws)), ("classic bn256", dict()), ("classic bn128", dict(block_n=128)), ("streamk bn128", dict(block_n=128, workspace=ws))]:
    out = torch.full((M, N), float("nan"), device="cuda")
    gm.gemm(M, N, K, a_hi, b_hi, a_lo, b_lo, b_mn=True, nterms=3, out_f32=out, act="mul_dgelu", aux=aux, **kw)
    torch.cuda.synchronize()
    d = (out.double() - ref64).abs()
    bad = d > 1e-3 * ref64.abs().max()
    print(name, "rel err", (d.norm() / ref64.norm()).item(), "bad", bad.sum().item())
    if bad.any():
        bn = 128 if "128" in name else 256
        tm = bad.view(25, 128, N // bn, bn).float().mean(dim=(1, 3))
        print((tm * 100).round().int()[:6])
        i, j = (tm > 0).nonzero()[0].tolist()
        blk = bad[i*128:(i+1)*128, j*bn:(j+1)*bn]
        print("  tile", i, j, "bad rows", blk.any(1).sum().item(), "bad cols", blk.any(0).nonzero().flatten().tolist()[:20], "...")
