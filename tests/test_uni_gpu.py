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


@pytest.mark.gpu
@pytest.mark.parametrize("hw", [(256, 256), (300, 300), (512, 512), (200, 200), (256, 320)])
def test_resize_is_bit_identical_to_pillow_restatement(hw):
    """`transforms.Resize(224)` on the GPU (csrc/preproc.cu) against oracle/resize_oracle.py, which is pinned to Pillow itself in
    tests/test_oracle_cpu.py: integer arithmetic, so every byte must match."""
    import numpy as np
    from oracle import resize_oracle as R
    from sequoia_pub_b200 import preproc
    rs = np.random.RandomState(hw[0])
    tiles = np.stack([(rs.rand(*hw, 3) * 255).astype(np.uint8),
                      np.clip(np.cumsum(rs.randn(*hw, 3), axis=1) * 8 + 128, 0, 255).astype(np.uint8),
                      np.full(hw + (3,), 255, np.uint8)])
    got = preproc.resize_tiles(torch.from_numpy(tiles).cuda(), 224).cpu().numpy()
    for i in range(len(tiles)):
        want = R.resize(tiles[i], 224)
        assert got[i].shape == want.shape and np.array_equal(got[i], want), (hw, i, int(np.abs(got[i].astype(int) - want.astype(int)).max()))


@pytest.mark.gpu
def test_uni_slide_through_the_pipeline_driver(tmp_path):
    """`--feat_type uni` of compute_features_hdf5.py:53-56,110-136 through pipeline.extract_slide: 256-px tiles from a patch file,
    Resize(224) + ToTensor + Normalize + ViT on the GPU, dataset "uni_features" [n, 1024]; rows must equal the extractor applied to
    the Pillow-restatement-resized tiles, in key order."""
    import os
    import numpy as np
    from oracle import resize_oracle as R
    from oracle import uni_oracle as U
    from sequoia_pub_b200 import hdf5, pipeline
    from sequoia_pub_b200.extract import SlideExtractor
    from sequoia_pub_b200.uni import VisionTransformer
    m = VisionTransformer(depth=2).eval()
    m.load_state_dict(U.make_state_dict(3, depth=2))
    m = m.cuda()
    rs = np.random.RandomState(8)
    tiles = {f"{256 * (i % 7)}_{256 * (i // 7)}": np.clip(rs.randint(0, 256, size=(1, 1, 3)) + rs.randint(-50, 50, size=(256, 256, 3)), 0, 255).astype(np.uint8)
             for i in range(23)}
    patch_file = tmp_path / "patches" / "S1" / "S1.hdf5"
    os.makedirs(patch_file.parent)
    with hdf5.File(patch_file, "w") as f:
        for k, v in tiles.items():
            f.create_dataset(k, data=v)
    feature_file = tmp_path / "features" / "P" / "S1" / "S1.h5"
    feats = pipeline.extract_slide(m, patch_file, feature_file, feat_type="uni", batch_size=8, prefer_h5py=False)
    assert feats.shape == (23, 1024) and feats.dtype == np.float32
    keys = sorted(tiles, key=str.encode)
    resized = torch.from_numpy(np.stack([R.resize(tiles[k], 224) for k in keys])).cuda()
    want = torch.cat([m.extract_uint8(resized[b:b + 8]) for b in range(0, 23, 8)]).cpu().numpy()
    assert np.array_equal(feats, want)
    with torch.no_grad():
        ref = U.forward(U.to_double(U.make_state_dict(3, depth=2)), U.preprocess(resized[:3].cpu()).double())
    assert _rel(torch.from_numpy(feats[:3]), ref) < TOL
    with hdf5.File(feature_file) as f:
        assert list(f.keys()) == ["uni_features"] and np.array_equal(f["uni_features"][:], feats)
    with pytest.raises(ValueError):                                   # a ResNet dataset name for a UNI model is refused
        pipeline.extract_slide(m, patch_file, tmp_path / "x.h5", feat_type="resnet", prefer_h5py=False)
    cf = pipeline.reduce_slide(feature_file, num_clusters=10, feat_name="uni_features", prefer_h5py=False)
    assert cf.shape == (10, 1024)
    # 224-px tiles skip the resize; an extractor object can be reused across slides
    ex = SlideExtractor(m, 8, (224, 224))
    assert np.array_equal(ex(resized.cpu()), want)
