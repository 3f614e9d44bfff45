"""UNI ViT-L/16 extractor: CUDA path vs the oracle restatement of timm's forward (PARITY UNPINNED: no timm / UNI weights
exist in this environment; see oracle/uni_oracle.py).  bf16 tensor-core operands, fp32 accumulation and residual stream:
north_star sets no bar for backbone features; tolerance 1e-2 L2-relative is written here."""
import pytest
import torch

TOL = 1e-2


def _rel(a, b):
    return ((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm()).item()


def test_state_dict_schema_matches_timm_names():
    from oracle import uni_oracle as U
    from sequoia_pub_b200.uni import create_model
    m = create_model("vit_large_patch16_224", img_size=224, patch_size=16, init_values=1e-5, num_classes=0, dynamic_img_size=True)
    want = U.param_shapes()
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == want
    m.load_state_dict(U.make_state_dict(0), strict=True)
    with pytest.raises(RuntimeError):
        m.eval()(torch.zeros(1, 3, 224, 224))         # no CPU fallback


@pytest.mark.gpu
@pytest.mark.parametrize("depth,batch", [(2, 3), (24, 2)])
def test_extract_matches_oracle(depth, batch):
    from oracle import uni_oracle as U
    from sequoia_pub_b200.uni import VisionTransformer
    sd = U.make_state_dict(1, depth=depth)
    m = VisionTransformer(depth=depth).eval()
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    patches = U.make_patches(5, batch)
    with torch.no_grad():
        want = U.forward(U.to_double(sd), U.preprocess(patches).double())
    got_u8 = m.extract_uint8(patches.cuda())
    got_f32 = m(U.preprocess(patches).cuda())
    e1, e2 = _rel(got_u8, want), _rel(got_f32, want)
    print(f"\n[uni parity] depth {depth}: L2-rel vs fp64 oracle: uint8 path {e1:.3e}, fp32 path {e2:.3e}")
    assert got_u8.shape == (batch, 1024) and e1 < TOL and e2 < TOL
    assert torch.equal(m.extract_uint8(patches.cuda()), got_u8)          # deterministic
    one = m.extract_uint8(patches[:1].cuda())
    assert _rel(one, got_u8[:1]) < 1e-3                                   # batch-size independent


@pytest.mark.gpu
def test_throughput_report(capsys):
    from oracle import uni_oracle as U
    from sequoia_pub_b200.uni import VisionTransformer
    m = VisionTransformer().eval()
    m.load_state_dict(U.make_state_dict(0))
    m = m.cuda()
    x = torch.randint(0, 256, (64, 224, 224, 3), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        m.extract_uint8(x)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(3):
        m.extract_uint8(x)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 3
    with capsys.disabled():
        print(f"\n[uni timing] batch 64: {ms:.2f} ms -> {64 / ms * 1e3:.0f} patches/s ({123.107e9 * 64 / ms / 1e9:.0f} TFLOP/s algorithmic)")


@pytest.mark.gpu
def test_two_lane_extraction_is_bit_identical():
    from oracle import uni_oracle as U
    from sequoia_pub_b200.uni import VisionTransformer
    m = VisionTransformer(depth=2).eval()
    m.load_state_dict(U.make_state_dict(2, depth=2))
    m = m.cuda()
    tiles = U.make_patches(9, 11).cuda()
    seq = torch.cat([m.extract_uint8(tiles[b:b + 4]) for b in range(0, 11, 4)])
    assert torch.equal(m.extract_many(tiles, batch_size=4, lanes=2), seq)
