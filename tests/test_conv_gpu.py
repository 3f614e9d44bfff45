"""CTA-pair tcgen05 convolution kernel (csrc/convgemm.cuh, `sq_conv_bf16`) against torch's fp64 conv2d on the same
bf16-rounded operands: out = relu(conv(x, w) + shift [+ residual]) rounded to bf16 (src/resnet.py:73-93 with BatchNorm folded).
Covers 1x1 / 3x3 / strided geometries of every ResNet-50 stage, the halo-staged 3x3 mode, both CTA-group sizes, every tile
width, ragged row counts (batch not a multiple of the tile) and the residual path."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _gm():
    from sequoia_pub_b200 import _gemm, _lib
    _lib.require_device()
    return _gemm


def _case(B, H, Cin, Cout, k, stride, res, relu, block_n=0, cta_group=0, seed=0):
    gm = _gm()
    g = torch.Generator(device="cuda").manual_seed(seed)
    pad = k // 2
    x = torch.randn(B, H, H, Cin, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(Cout, k, k, Cin, device="cuda", generator=g) * (1.0 / (k * k * Cin) ** 0.5)).to(torch.bfloat16)
    sh = torch.randn(Cout, device="cuda", generator=g)
    Ho = (H + 2 * pad - k) // stride + 1
    r = torch.randn(B, Ho, Ho, Cout, device="cuda", generator=g).to(torch.bfloat16) if res else None
    out = torch.full((B, Ho, Ho, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    gm.conv_bf16(x, w, sh, r, relu, stride, pad, block_n, cta_group, out=out)
    torch.cuda.synchronize()
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double().permute(0, 3, 1, 2), stride=stride, padding=pad).permute(0, 2, 3, 1)
    ref = ref + sh.double()
    if res:
        ref = ref + r.double()
    if relu:
        ref = ref.clamp_min(0)
    assert torch.isfinite(out.float()).all()
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 6e-3, err            # one bf16 rounding of the result (2^-8 relative) plus fp32 accumulation noise
    return out


@pytest.mark.parametrize("geom", [
    (4, 64, 64, 64, 1, 1, False),      # layer1 conv1 (first block)
    (4, 64, 64, 64, 3, 1, False),      # layer1 conv2: halo mode, 64-wide tiles
    (4, 64, 64, 256, 1, 1, True),      # layer1 conv3 + residual
    (4, 64, 256, 64, 1, 1, False),     # layer1 conv1 (later blocks)
    (4, 64, 128, 128, 3, 2, False),    # layer2 conv2, stride 2 (per-tap boxes)
    (4, 64, 256, 512, 1, 2, False),    # layer2 downsample, 1x1 stride 2
    (4, 32, 128, 128, 3, 1, False),    # layer2 conv2: halo mode, two channel blocks
    (4, 32, 128, 512, 1, 1, True),
    (8, 16, 256, 256, 3, 1, False),    # layer3 conv2
    (8, 16, 256, 1024, 1, 1, True),
    (8, 16, 1024, 256, 1, 1, False),
    (16, 8, 512, 512, 3, 1, False),    # layer4 conv2: two images per 128-row tile
    (16, 8, 512, 2048, 1, 1, True),
])
def test_resnet_geometries(geom):
    B, H, Cin, Cout, k, stride, res = geom
    _case(B, H, Cin, Cout, k, stride, res, True)


@pytest.mark.parametrize("cta_group", [1, 2])
@pytest.mark.parametrize("block_n", [64, 128, 256])
def test_tile_widths_and_cta_groups(block_n, cta_group):
    if cta_group == 1 and block_n == 256:
        pytest.skip("a single CTA has no 256-wide configuration (falls back to 128)")
    _case(3, 32, 128, 256, 1, 1, True, True, block_n, cta_group)       # 3072 rows: 24 tiles, ragged pair count for some grids
    _case(3, 32, 64, 256, 3, 1, False, True, block_n, cta_group)


def test_ragged_rows_and_no_relu():
    _case(5, 16, 64, 64, 1, 1, False, False)      # 1280 rows = 10 tiles; 5 CTA pairs
    _case(1, 16, 64, 128, 3, 1, False, True)      # 256 rows: a single pair
    _case(3, 8, 128, 128, 1, 1, True, False)      # 192 rows: the second tile of the pair is half out of range
    _case(7, 8, 64, 64, 3, 1, False, True)        # 448 rows, two images per tile, odd image count


def test_fused_average_pool_matches_unfused():
    """Last convolution + AvgPool2d(7) on the 8x8 map (top-left 7x7, src/resnet.py:110,166) fused into the epilogue: the
    extractor output must agree with the un-fused path (fp32 map + pooling kernel) to fp32 summation-order noise."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, torch; sys.path.insert(0, %r)\n"
            "from oracle import resnet50_oracle as O\n"
            "from sequoia_pub_b200.resnet import resnet50\n"
            "m = resnet50().eval(); m.load_state_dict(O.make_state_dict(0)); m = m.cuda()\n"
            "f = m.extract_uint8(O.make_patches(5, 3).cuda()); torch.save(f.cpu(), sys.argv[1])\n") % root
    outs = []
    for fused in ("1", "0"):
        path = f"/tmp/sq_pool_{fused}.pt"
        subprocess.run([sys.executable, "-c", code, path], env=dict(os.environ, SQ_POOL_FUSED=fused), check=True, timeout=300)
        outs.append(torch.load(path))
    rel = ((outs[0] - outs[1]).norm() / outs[1].norm()).item()
    assert rel < 1e-6, rel


@pytest.mark.parametrize("B,H,W", [(2, 64, 64), (1, 16, 8), (3, 32, 24), (5, 64, 64)])
def test_fused_layer1_tail_is_bit_identical_to_two_launches(B, H, W):
    """`sq_bneck_l1_bf16` (csrc/fusedconv.cuh: conv3x3 64->64 + ReLU feeding conv1x1 64->256 + residual + ReLU through tensor memory)
    against the same two convolutions as separate `sq_conv_bf16` launches (bit-identical: the intermediate is rounded to bf16 either
    way) and against torch fp64 on the bf16-rounded operands (src/resnet.py:73-93)."""
    gm = _gm()
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + H + W)
    x = torch.randn(B, H, W, 64, device="cuda", generator=g).relu().to(torch.bfloat16)
    w2 = (torch.randn(64, 3, 3, 64, device="cuda", generator=g) * (1.0 / 576 ** 0.5)).to(torch.bfloat16)
    w3 = (torch.randn(256, 1, 1, 64, device="cuda", generator=g) * (1.0 / 8.0)).to(torch.bfloat16)
    s2 = torch.randn(64, device="cuda", generator=g) * 0.5
    s3 = torch.randn(256, device="cuda", generator=g) * 0.5
    r = torch.randn(B, H, W, 256, device="cuda", generator=g).to(torch.bfloat16)
    mid = gm.conv_bf16(x, w2, s2, None, True, 1, 1)
    want = gm.conv_bf16(mid, w3, s3, r, True, 1, 0)
    out = torch.full((B, H, W, 256), float("nan"), device="cuda", dtype=torch.bfloat16)
    gm.bneck_l1_bf16(x, w2, s2, w3, s3, r, out=out)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    assert torch.equal(out, want)
    mid64 = (F.conv2d(x.double().permute(0, 3, 1, 2), w2.double().permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1) + s2.double()).clamp_min(0)
    mid64 = mid64.to(torch.bfloat16).double()              # the intermediate is a bf16 tensor in the reference pipeline of this package too
    ref = (mid64 @ w3.double().reshape(256, 64).t() + s3.double() + r.double()).clamp_min(0)
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 8e-3, err


def test_fused_layer1_tail_rejects_untiled_maps():
    gm = _gm()
    x = torch.zeros(1, 56, 56, 64, device="cuda", dtype=torch.bfloat16)
    w2 = torch.zeros(64, 3, 3, 64, device="cuda", dtype=torch.bfloat16); w3 = torch.zeros(256, 64, device="cuda", dtype=torch.bfloat16)
    s2 = torch.zeros(64, device="cuda"); s3 = torch.zeros(256, device="cuda")
    r = torch.zeros(1, 56, 56, 256, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        gm.bneck_l1_bf16(x, w2, s2, w3, s3, r)


@pytest.mark.parametrize("B,H,W", [(2, 64, 64), (1, 16, 8), (3, 32, 24)])
def test_fused_layer1_first_block_with_downsample_matches_reference(B, H, W):
    """`sq_bneck_l1_ds_bf16`: the downsample branch (1x1, 64 -> 256 on the block input) is computed inside the fused tail, so the residual sum
    is formed in fp32 and rounded once.  Reference: torch fp64 on the bf16-rounded operands with the bf16 intermediate (src/resnet.py:73-93,
    downsample = conv1x1 + bn without ReLU); it must also agree with the two-launch path up to the extra rounding of the residual."""
    gm = _gm()
    g = torch.Generator(device="cuda").manual_seed(7000 + B * 100 + H + W)
    x = torch.randn(B, H, W, 64, device="cuda", generator=g).relu().to(torch.bfloat16)            # block input
    x1 = torch.randn(B, H, W, 64, device="cuda", generator=g).relu().to(torch.bfloat16)           # conv1's output
    w2 = (torch.randn(64, 3, 3, 64, device="cuda", generator=g) * (1.0 / 576 ** 0.5)).to(torch.bfloat16)
    w3 = (torch.randn(256, 1, 1, 64, device="cuda", generator=g) * (1.0 / 8.0)).to(torch.bfloat16)
    wds = (torch.randn(256, 1, 1, 64, device="cuda", generator=g) * (1.0 / 8.0)).to(torch.bfloat16)
    s2 = torch.randn(64, device="cuda", generator=g) * 0.5
    s3 = torch.randn(256, device="cuda", generator=g) * 0.5
    sds = torch.randn(256, device="cuda", generator=g) * 0.5
    out = torch.full((B, H, W, 256), float("nan"), device="cuda", dtype=torch.bfloat16)
    gm.bneck_l1_ds_bf16(x1, w2, s2, w3, s3, x, wds, sds, out=out)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    mid = (F.conv2d(x1.double().permute(0, 3, 1, 2), w2.double().permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1) + s2.double()).clamp_min(0)
    mid = mid.to(torch.bfloat16).double()
    ref = (mid @ w3.double().reshape(256, 64).t() + s3.double() + x.double() @ wds.double().reshape(256, 64).t() + sds.double()).clamp_min(0)
    scale = ref.abs().max().item()
    assert (out.double() - ref).abs().max().item() / scale < 6e-3
    res = gm.conv_bf16(x, wds, sds, None, False, 1, 0)                                            # two-launch path: residual rounded to bf16 first
    two = gm.bneck_l1_bf16(x1, w2, s2, w3, s3, res)
    assert (out.double() - two.double()).abs().max().item() / scale < 1.2e-2

