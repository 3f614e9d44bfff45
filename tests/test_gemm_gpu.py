"""tcgen05 GEMM / implicit-GEMM conv kernel vs plain PyTorch (fp32 / fp64) on the same inputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _mods():
    from sequoia_pub_b200 import _gemm, _lib
    _lib.require_device()
    return _gemm


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _mk(rows, cols, seed, mn_major=False):
    """Random fp32 matrix [rows(MN), cols(K)] and the stored tensor in the requested major."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(rows, cols, device="cuda", generator=g)
    return x


def _store(x, mn_major):
    # logical [MN, K]; MN-major storage is the transposed matrix [K, MN], row-major. ld padded to 8.
    t = x.t().contiguous() if mn_major else x.contiguous()
    r, c = t.shape
    ld = (c + 7) // 8 * 8
    buf = torch.zeros(r, ld, device="cuda")
    buf[:, :c] = t
    return buf[:, :c]


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("shape", [(256, 256, 256), (300, 200, 100), (128, 64, 64), (3200, 1024, 512)])
@pytest.mark.parametrize("block_n", [64, 128, 256])
def test_plain_bf16(a_mn, b_mn, shape, block_n):
    gm = _mods()
    M, N, K = shape
    A = _mk(M, K, 1); B = _mk(N, K, 2)
    As, Bs = _store(A, a_mn), _store(B, b_mn)
    a_hi, _ = gm.split_planes(As, want_lo=False)
    b_hi, _ = gm.split_planes(Bs, want_lo=False)
    out = torch.full((M, N), float("nan"), device="cuda")
    gm.gemm(M, N, K, a_hi, b_hi, a_mn=a_mn, b_mn=b_mn, out_f32=out, block_n=block_n)
    torch.cuda.synchronize()
    Ah = (a_hi[:, :As.shape[1]].float().t() if a_mn else a_hi[:, :K].float())
    Bh = (b_hi[:, :Bs.shape[1]].float().t() if b_mn else b_hi[:, :K].float())
    ref = Ah.double() @ Bh.double().t()
    assert torch.isfinite(out).all()
    assert _rel(out, ref) < 2e-6


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (True, True), (False, True)])
@pytest.mark.parametrize("shape", [(3200, 2048, 2048), (32, 1000, 2048), (2048, 2048, 3200), (100, 20530, 32)])
def test_split3_matches_fp32(a_mn, b_mn, shape):
    gm = _mods()
    M, N, K = shape
    A = _mk(M, K, 3); B = _mk(N, K, 4)
    As, Bs = _store(A, a_mn), _store(B, b_mn)
    a_hi, a_lo = gm.split_planes(As)
    b_hi, b_lo = gm.split_planes(Bs)
    ldc = (N + 3) // 4 * 4
    out = torch.full((M, ldc), float("nan"), device="cuda")[:, :N]
    bias = torch.randn(N, device="cuda")
    gm.gemm(M, N, K, a_hi, b_hi, a_lo, b_lo, a_mn=a_mn, b_mn=b_mn, nterms=3, out_f32=out, bias=bias)
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t() + bias.double()
    assert torch.isfinite(out).all()
    assert _rel(out, ref) < 2e-5


@pytest.mark.parametrize("split_k", [2, 5])
def test_split_k(split_k):
    gm = _mods()
    M, N, K = 32, 2048, 4096
    A = _mk(M, K, 5); B = _mk(N, K, 6)
    a_hi, a_lo = gm.split_planes(A); b_hi, b_lo = gm.split_planes(B)
    out = torch.full((M, N), float("nan"), device="cuda")
    bias = torch.randn(N, device="cuda")
    gm.gemm(M, N, K, a_hi, b_hi, a_lo, b_lo, nterms=3, out_f32=out, bias=bias, act="gelu", split_k=split_k)
    torch.cuda.synchronize()
    ref = F.gelu(A.double() @ B.double().t() + bias.double())
    assert _rel(out, ref) < 2e-5


def test_epilogues():
    gm = _mods()
    M, N, K = 400, 1024, 256
    A = _mk(M, K, 7); B = _mk(N, K, 8) * 0.1
    a_hi, a_lo = gm.split_planes(A); b_hi, b_lo = gm.split_planes(B)
    base = A.double() @ B.double().t()
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    # bias + residual + relu, planes out
    out = torch.empty(M, N, device="cuda")
    ohi = torch.empty(M, N, device="cuda", dtype=torch.bfloat16); olo = torch.empty_like(ohi)
    pre = torch.empty(M, N, device="cuda")
    gm.gemm(M, N, K, a_hi, b_hi, a_lo, b_lo, nterms=3, out_f32=out, out_hi=ohi, out_lo=olo, bias=bias, res_f32=res,
            save_pre=pre, act="relu")
    torch.cuda.synchronize()
    ref_pre = base + bias.double() + res.double()
    assert _rel(pre, ref_pre) < 2e-5
    assert _rel(out, ref_pre.clamp_min(0)) < 2e-5
    assert _rel(ohi.float() + olo.float(), out) < 1e-4 ** 1  # hi+lo reproduces fp32 to ~2^-16
    assert torch.equal(ohi, out.to(torch.bfloat16))
    # bf16 residual
    resb = res.to(torch.bfloat16)
    gm.gemm(M, N, K, a_hi, b_hi, a_lo, b_lo, nterms=3, out_f32=out, res_bf=resb, act="none")
    torch.cuda.synchronize()
    assert _rel(out, base + resb.double()) < 2e-5
    # per-head LN(64) + GELU
    gam = torch.rand(N, device="cuda") + 0.5; bet = torch.randn(N, device="cuda") * 0.1
    gm.gemm(M, N, K, a_hi, b_hi, a_lo, b_lo, nterms=3, out_f32=out, bias=bias, save_pre=pre, ln_gamma=gam, ln_beta=bet,
            act="ln64_gelu")
    torch.cuda.synchronize()
    v = (base + bias.double()).view(M, N // 64, 64)
    ln = F.layer_norm(v, (64,), eps=1e-5).view(M, N) * gam.double() + bet.double()
    assert _rel(pre, base + bias.double()) < 2e-5
    assert _rel(out, F.gelu(ln)) < 5e-5
    # row bias (per group of 100 rows) + gelu
    rb = torch.randn(4, N, device="cuda")
    gm.gemm(M, N, K, a_hi, b_hi, a_lo, b_lo, nterms=3, out_f32=out, rowbias=rb, rowbias_div=100, act="gelu")
    torch.cuda.synchronize()
    assert _rel(out, F.gelu(base + rb.double().repeat_interleave(100, 0))) < 2e-5
    # multiply by gelu'(aux)
    aux = torch.randn(M, N, device="cuda")
    gm.gemm(M, N, K, a_hi, b_hi, a_lo, b_lo, nterms=3, out_f32=out, aux=aux, act="mul_dgelu", alpha=0.5)
    torch.cuda.synchronize()
    ad = aux.double().requires_grad_(True)
    F.gelu(ad).sum().backward()
    assert _rel(out, 0.5 * base * ad.grad) < 2e-5


def test_block_diagonal():
    gm = _mods()
    M, H, d = 300, 16, 64
    A = _mk(M, H * d, 9); W = _mk(H * d, d, 10)   # W[h*64 + j, i]
    a_hi, a_lo = gm.split_planes(A); w_hi, w_lo = gm.split_planes(W)
    out = torch.empty(M, H * d, device="cuda")
    gm.gemm(M, H * d, d, a_hi, w_hi, a_lo, w_lo, nterms=3, out_f32=out, block_n=64, a_koff_per_ntile=64)
    torch.cuda.synchronize()
    ref = torch.einsum("mhi,hji->mhj", A.double().view(M, H, d), W.double().view(H, d, d)).reshape(M, H * d)
    assert _rel(out, ref) < 2e-5


CONVS = [  # (batch, H, W, C, Cout, R, stride, pad)
    (2, 64, 64, 64, 64, 1, 1, 0),
    (2, 64, 64, 64, 64, 3, 1, 1),
    (2, 64, 64, 128, 128, 3, 2, 1),
    (3, 64, 64, 256, 512, 1, 2, 0),
    (2, 32, 32, 128, 128, 3, 1, 1),
    (2, 32, 32, 256, 256, 3, 2, 1),
    (3, 16, 16, 256, 256, 3, 1, 1),
    (4, 16, 16, 512, 512, 3, 2, 1),
    (5, 8, 8, 512, 512, 3, 1, 1),
    (3, 8, 8, 2048, 512, 1, 1, 0),
    (1, 128, 128, 64, 64, 1, 1, 0),
]


@pytest.mark.parametrize("cfg", CONVS)
def test_conv(cfg):
    gm = _mods()
    batch, H, W, Cin, Cout, R, stride, pad = cfg
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(batch, H, W, Cin, device="cuda", generator=g).to(torch.bfloat16)          # NHWC
    w = (torch.randn(Cout, R, R, Cin, device="cuda", generator=g) * 0.05).to(torch.bfloat16)   # OHWI
    Ho = (H + 2 * pad - R) // stride + 1
    Wo = (W + 2 * pad - R) // stride + 1
    M = batch * Ho * Wo
    shift = torch.randn(Cout, device="cuda")
    res = torch.randn(M, Cout, device="cuda").to(torch.bfloat16)
    out = torch.empty(M, Cout, device="cuda", dtype=torch.bfloat16)
    gm.gemm(M, Cout, R * R * Cin, x.view(-1, Cin), w.view(Cout, -1), conv=(batch, H, W, Cin, Ho, Wo, R, R, stride, pad),
            out_hi=out, bias=shift, res_bf=res, act="relu")
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), stride=stride, padding=pad)
    ref = (ref.permute(0, 2, 3, 1).reshape(M, Cout) + shift + res.float()).clamp_min(0)
    err = _rel(out.float(), ref)
    assert err < 6e-3, err   # bf16 output rounding


def test_timing_report(capsys):
    """Not a pass/fail test: prints achieved TFLOP/s for a few shapes so the first GPU run gives a perf signal."""
    gm = _mods()
    for (M, N, K, nt) in [(8192, 8192, 8192, 1), (3200, 2048, 2048, 3), (3200, 2048, 2048, 1), (262144, 256, 64, 1)]:
        A = _mk(M, K, 1); B = _mk(N, K, 2)
        a_hi, a_lo = gm.split_planes(A); b_hi, b_lo = gm.split_planes(B)
        out = torch.empty(M, N, device="cuda")
        for bn in (128, 256):
            for _ in range(3):
                gm.gemm(M, N, K, a_hi, b_hi, a_lo, b_lo, nterms=nt, out_f32=out, block_n=bn)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(10):
                gm.gemm(M, N, K, a_hi, b_hi, a_lo, b_lo, nterms=nt, out_f32=out, block_n=bn)
            e.record(); torch.cuda.synchronize()
            ms = s.elapsed_time(e) / 10
            with capsys.disabled():
                print(f"\n[gemm timing] M={M} N={N} K={K} terms={nt} BN={bn}: {ms*1e3:.1f} us, "
                      f"{2.0*M*N*K*nt/ms/1e9:.1f} TFLOP/s (bf16 MMA work)")


def test_block_diagonal_backward_modes():
    """Per-head dgrad (B read transposed from the [H*64, 128] c.weight layout) and diagonal-block wgrad."""
    gm = _mods()
    M, H, d = 300, 16, 64
    dC = _mk(M, H * d, 21)                                   # dL/dCpre [tokens, H*64]
    Wc = _mk(H * d, 2 * d, 22) * 0.2                         # c.weight of all heads, [H*64, 128]
    local = _mk(M, H * d, 23)
    a_hi, a_lo = gm.split_planes(dC); w_hi, w_lo = gm.split_planes(Wc); l_hi, l_lo = gm.split_planes(local)
    for half in (0, 1):                                      # Wc[:, :64] and Wc[:, 64:]
        out = torch.full((M, H * d), float("nan"), device="cuda")
        gm.gemm(M, H * d, d, a_hi, w_hi[:, half * 64:], a_lo, w_lo[:, half * 64:], b_mn=True, ldb=128, nterms=3, out_f32=out,
                block_n=64, a_koff_per_ntile=64, b_koff_per_ntile=64, b_nadj_per_ntile=-64, b_map_mn=64, b_map_k=H * d)
        torch.cuda.synchronize()
        W = Wc.double().view(H, d, 2 * d)[:, :, half * 64:(half + 1) * 64]          # [h, j, i]
        ref = torch.einsum("mhj,hji->mhi", dC.double().view(M, H, d), W).reshape(M, H * d)
        assert _rel(out, ref) < 2e-5
    dW = torch.full((H * d, 2 * d), float("nan"), device="cuda")
    gm.gemm(H * d, H * d, M, a_hi, l_hi, a_lo, l_lo, a_mn=True, b_mn=True, nterms=3, out_f32=dW, diag64=1, block_n=64)
    torch.cuda.synchronize()
    ref = torch.einsum("mhj,mhi->hji", dC.double().view(M, H, d), local.double().view(M, H, d)).reshape(H * d, d)
    assert _rel(dW[:, :64], ref) < 2e-5
    assert torch.isnan(dW[:, 64:]).all()                      # the other half of c.weight's gradient is untouched


@pytest.mark.parametrize("case", [(3200, 2048, 2048, 3, False, False, 0), (3200, 2048, 2048, 3, True, True, 0), (1000, 1024, 1536, 1, False, False, 0),
                                  (3200, 1024, 2048, 3, False, False, 128), (2500, 2304, 1024, 1, False, True, 256), (32, 20530, 2048, 3, False, False, 0)])
def test_stream_k_scheduling(case):
    """GEMMs whose tile count leaves the last wave mostly empty are scheduled stream-K (equal k-block shares per CTA,
    partial tiles exchanged through the workspace); results must not depend on it."""
    gm = _mods()
    M, N, K, nt, a_mn, b_mn, bn = case
    A = _mk(M, K, 31); B = _mk(N, K, 32)
    As, Bs = _store(A, a_mn), _store(B, b_mn)
    a_hi, a_lo = gm.split_planes(As); b_hi, b_lo = gm.split_planes(Bs)
    ws = torch.empty(160 * 256 * 128 + 4096, dtype=torch.float32, device="cuda")
    bias = torch.randn(N, device="cuda")
    ldc = (N + 3) // 4 * 4
    out = torch.full((M, ldc), float("nan"), device="cuda")[:, :N]
    ref_out = torch.full((M, ldc), float("nan"), device="cuda")[:, :N]
    kw = dict(a_mn=a_mn, b_mn=b_mn, nterms=nt, bias=bias, act="gelu", block_n=bn)
    gm.gemm(M, N, K, a_hi, b_hi, a_lo if nt == 3 else None, b_lo if nt == 3 else None, out_f32=out, workspace=ws, **kw)
    gm.gemm(M, N, K, a_hi, b_hi, a_lo if nt == 3 else None, b_lo if nt == 3 else None, out_f32=ref_out, **kw)      # classic scheduling
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    if nt == 3:
        ref = F.gelu(A.double() @ B.double().t() + bias.double())
        assert _rel(out, ref) < 2e-5
    assert _rel(out, ref_out) < 5e-6          # same products, different (fixed) fp32 summation grouping
    again = torch.empty_like(out)
    gm.gemm(M, N, K, a_hi, b_hi, a_lo if nt == 3 else None, b_lo if nt == 3 else None, out_f32=again, workspace=ws, **kw)
    torch.cuda.synchronize()
    assert torch.equal(again, out)            # deterministic


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True)])
def test_split3_auto_tile_width_192(a_mn, b_mn):
    """M=3200, N=2048 picks 192-wide tiles (275 tiles = 2 waves on 148 SMs); last n-tile is 128 columns wide."""
    gm = _mods()
    M, N, K = 3200, 2048, 1024
    A = _mk(M, K, 41); B = _mk(N, K, 42)
    As, Bs = _store(A, a_mn), _store(B, b_mn)
    a_hi, a_lo = gm.split_planes(As); b_hi, b_lo = gm.split_planes(Bs)
    bias = torch.randn(N, device="cuda"); res = torch.randn(M, N, device="cuda")
    out = torch.full((M, N), float("nan"), device="cuda"); ohi = torch.empty(M, N, device="cuda", dtype=torch.bfloat16); olo = torch.empty_like(ohi)
    gm.gemm(M, N, K, a_hi, b_hi, a_lo, b_lo, a_mn=a_mn, b_mn=b_mn, nterms=3, out_f32=out, out_hi=ohi, out_lo=olo, bias=bias, res_f32=res)
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t() + bias.double() + res.double()
    assert _rel(out, ref) < 2e-5
    assert _rel(ohi.float() + olo.float(), out) < 1e-4
