"""A/B of the CTA-pair / TMA-epilogue convolution kernel (csrc/convgemm.cuh) against the round-1 kernel (SQ_CONVGEMM=0):
features must be bit-identical (same fp32 accumulation order, same rounding points); prints ms per batch of 64 for both.

    python tools/convgemm_check.py [extra env assignments for arm 1, e.g. SQ_CONV_CG=1]
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARM = r'''
import sys, torch
sys.path.insert(0, %r)
from oracle import resnet50_oracle as O
from sequoia_pub_b200.resnet import resnet50
m = resnet50().eval(); m.load_state_dict(O.make_state_dict(0)); m = m.cuda()
x = torch.randint(0, 256, (64, 256, 256, 3), dtype=torch.uint8, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
f = m.extract_uint8(x)
f5 = m.extract_uint8(x[:5])
for _ in range(3): m.extract_uint8(x)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10): m.extract_uint8(x)
e.record(); torch.cuda.synchronize()
torch.save((f.cpu(), f5.cpu()), sys.argv[1])
big = torch.randint(0, 256, (1024, 256, 256, 3), dtype=torch.uint8, device="cuda", generator=torch.Generator(device="cuda").manual_seed(4))
import os
res = []
for lanes in (2, 3, 4):
    m.extract_many(big, lanes=lanes)
    s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s2.record(); m.extract_many(big, lanes=lanes); e2.record(); torch.cuda.synchronize()
    res.append(round(1024 / s2.elapsed_time(e2), 2))
print("ms per batch of 64:", round(s.elapsed_time(e) / 10, 4), " 2/3/4 lanes:", res, "k patches/s")
''' % ROOT

if __name__ == "__main__":
    import torch
    extra = dict(a.split("=", 1) for a in sys.argv[1:] if a != "--skip-old")
    outs = []
    arms = (("old", {"SQ_CONVGEMM": "0"}), ("new", dict({"SQ_CONVGEMM": "1"}, **extra)))
    if "--skip-old" in sys.argv:
        arms = arms[1:]
    for arm, env_add in arms:
        path = f"/tmp/convgemm_{arm}.pt"
        r = subprocess.run([sys.executable, "-c", ARM, path], env=dict(os.environ, **env_add), capture_output=True, text=True, timeout=300)
        print(f"{arm} {env_add}: rc={r.returncode} {r.stdout.strip()} {r.stderr.strip()[-600:]}", flush=True)
        outs.append(torch.load(path) if r.returncode == 0 else None)
    if len(outs) == 2 and outs[0] is not None and outs[1] is not None:
        for k, name in ((0, "batch 64"), (1, "batch 5")):
            a, b = outs[0][k], outs[1][k]
            print(name, "bit-identical:", bool(torch.equal(a, b)), "max abs diff", float((a - b).abs().max()), "rel", float((a - b).norm() / a.norm()))
