"""ORACLE (test infrastructure, not product code) — CPU restatement of the reference ResNet-50 feature path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

Restates, as plain functional PyTorch fp32 on the CPU:
  * the preprocessing of pre_processing/compute_features_hdf5.py:49-51,119-120
    (uint8 HWC -> permute -> /255 -> Normalize(mean, std));
  * ResNet.forward_extract of src/resnet.py:155-170 (stem :156-159, Bottleneck :73-93 with the stride on the
    3x3 conv :64, downsample :123-128, AvgPool2d(7) :110,166 — at 256 px that is the top-left 7x7 of the 8x8 map).
Pinned against the reference class itself: tests/golden/gen_golden.py imports /root/reference/src/resnet.py,
loads the weights generated here and stores its outputs in tests/golden/resnet50_golden.npz
(tests/test_oracle_cpu.py checks this restatement against that file).
"""
import torch
import torch.nn.functional as F

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)
STAGES = ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2))


def conv_bn_names():
    """[(conv_prefix, bn_prefix, cin, cout, k, stride, pad)] in module order (src/resnet.py:101-109)."""
    out = [("conv1", "bn1", 3, 64, 7, 2, 3)]
    inpl = 64
    for li, (planes, blocks, stride) in enumerate(STAGES, start=1):
        for b in range(blocks):
            s = stride if b == 0 else 1
            p = f"layer{li}.{b}"
            out.append((f"{p}.conv1", f"{p}.bn1", inpl, planes, 1, 1, 0))
            out.append((f"{p}.conv2", f"{p}.bn2", planes, planes, 3, s, 1))
            out.append((f"{p}.conv3", f"{p}.bn3", planes, planes * 4, 1, 1, 0))
            if b == 0:
                out.append((f"{p}.downsample.0", f"{p}.downsample.1", inpl, planes * 4, 1, s, 0))
            inpl = planes * 4
    return out


def make_state_dict(seed=0):
    """Deterministic weights with NON-trivial BN statistics (default 0/1 stats would hide BN bugs)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for conv, bn, cin, cout, k, _, _ in conv_bn_names():
        fan = k * k * cout
        sd[f"{conv}.weight"] = torch.randn(cout, cin, k, k, generator=g) * (2.0 / fan) ** 0.5
        sd[f"{bn}.weight"] = torch.rand(cout, generator=g) * 0.5 + 0.25
        sd[f"{bn}.bias"] = torch.randn(cout, generator=g) * 0.1
        sd[f"{bn}.running_mean"] = torch.randn(cout, generator=g) * 0.1
        sd[f"{bn}.running_var"] = torch.rand(cout, generator=g) + 0.5
        sd[f"{bn}.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    sd["fc.weight"] = torch.randn(1000, 2048, generator=g) * 0.02
    sd["fc.bias"] = torch.zeros(1000)
    return sd


def make_patches(seed, n, size=256):
    """Synthetic uint8 tiles [n,size,size,3] (BASELINE config 2: randint(0,256) seeded by slide id)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (n, size, size, 3), generator=g, dtype=torch.uint8)


def preprocess(patches_u8):
    """uint8 [B,H,W,3] -> float32 [B,3,H,W]  (compute_features_hdf5.py:49-51,119-120)."""
    x = patches_u8.permute(0, 3, 1, 2).to(torch.float32) / 255.0
    mean = torch.tensor(MEAN, dtype=torch.float32).view(1, 3, 1, 1)
    std = torch.tensor(STD, dtype=torch.float32).view(1, 3, 1, 1)
    return (x - mean) / std


def _bn(x, sd, p, eps=1e-5):
    return F.batch_norm(x, sd[f"{p}.running_mean"], sd[f"{p}.running_var"], sd[f"{p}.weight"], sd[f"{p}.bias"], False, 0.0, eps)


def forward_extract(sd, x):
    """x float [B,3,H,W] -> [B,2048]; works in whatever dtype sd / x are (fp32 or fp64)."""
    x = F.relu(_bn(F.conv2d(x, sd["conv1.weight"], stride=2, padding=3), sd, "bn1"))
    x = F.max_pool2d(x, 3, stride=2, padding=1)
    for li, (planes, blocks, stride) in enumerate(STAGES, start=1):
        for b in range(blocks):
            p = f"layer{li}.{b}"
            s = stride if b == 0 else 1
            out = F.relu(_bn(F.conv2d(x, sd[f"{p}.conv1.weight"]), sd, f"{p}.bn1"))
            out = F.relu(_bn(F.conv2d(out, sd[f"{p}.conv2.weight"], stride=s, padding=1), sd, f"{p}.bn2"))
            out = _bn(F.conv2d(out, sd[f"{p}.conv3.weight"]), sd, f"{p}.bn3")
            res = x
            if b == 0:
                res = _bn(F.conv2d(x, sd[f"{p}.downsample.0.weight"], stride=s), sd, f"{p}.downsample.1")
            x = F.relu(out + res)
    x = F.avg_pool2d(x, 7)           # stride = kernel = 7: 8x8 -> 1x1 over rows/cols 0..6
    return x.reshape(x.shape[0], -1)


def to_double(sd):
    return {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
