"""A/B of the fused layer-1 bottleneck tail (SQ_BNECK_FUSE=1, default) against the two separate launches (SQ_BNECK_FUSE=0):
features must be bit-identical without the in-kernel downsample (SQ_BNECK_DS=0); batch time with 1 and 2 extractor lanes.  Each setting runs in its own process (the switch is read once)."""
import hashlib, os, subprocess, sys

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import resnet50_oracle as O
    from sequoia_pub_b200.resnet import resnet50
    m = resnet50().eval(); m.load_state_dict(O.make_state_dict(0)); m = m.cuda()
    x = torch.randint(0, 256, (1024, 256, 256, 3), dtype=torch.uint8, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    ref = torch.cat([m.extract_uint8(x[b:b + 64]) for b in range(0, 1024, 64)])
    small = m.extract_uint8(x[:5])
    torch.cuda.synchronize()
    assert torch.equal(small, ref[:5])
    print("sha", hashlib.sha256(ref.cpu().numpy().tobytes()).hexdigest()[:16], "finite", bool(torch.isfinite(ref).all()))
    for lanes in (1, 2):
        out = m.extract_many(x, lanes=lanes); torch.cuda.synchronize()
        assert torch.equal(out, ref), lanes
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(3): m.extract_many(x, out=out, lanes=lanes)
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 3
        print(f"lanes={lanes}: {ms / 16:.3f} ms/batch -> {1024 / ms * 1e3:.0f} patches/s")
else:
    # FUSE=1 DS=0 must reproduce FUSE=0 bit for bit; DS=1 (downsample inside the first block's tail) rounds the residual sum once
    # instead of twice, so its features differ in the last bits (parity against the reference: tests/test_resnet_gpu.py)
    for rep in range(int(os.environ.get("FUSE_AB_REPS", "2"))):
        for fuse, ds in (("0", "0"), ("1", "0"), ("1", "1")):
            env = dict(os.environ, SQ_BNECK_FUSE=fuse, SQ_BNECK_DS=ds)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True, timeout=150)
            print(f"== SQ_BNECK_FUSE={fuse} SQ_BNECK_DS={ds} rc={r.returncode}")
            print((r.stdout + r.stderr[-1500:]).strip())
