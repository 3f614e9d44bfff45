"""A/B driver for the opt-in pipelined convolution epilogue (csrc/gemm.cuh: EPI_CONV_PF, SQ_CONV_EPI_PF=1).  The switch is read
once per process, so each arm runs in its own interpreter:

    python tools/conv_pf_check.py            # runs both arms, compares the features bit for bit, prints ms per batch of 64

Not part of the test-suite: the variant was written after the round-1 GPU budget was spent and has not run on hardware yet."""
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
for _ in range(3): m.extract_uint8(x)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10): m.extract_uint8(x)
e.record(); torch.cuda.synchronize()
torch.save(f.cpu(), sys.argv[1])
print("ms per batch of 64:", s.elapsed_time(e) / 10)
''' % ROOT

if __name__ == "__main__":
    import torch
    outs = []
    for pf in ("0", "1"):
        path = f"/tmp/conv_pf_{pf}.pt"
        env = dict(os.environ, SQ_CONV_EPI_PF=pf)
        r = subprocess.run([sys.executable, "-c", ARM, path], env=env, capture_output=True, text=True, timeout=300)
        print(f"SQ_CONV_EPI_PF={pf}: rc={r.returncode} {r.stdout.strip()} {r.stderr.strip()[-300:]}")
        outs.append(torch.load(path) if r.returncode == 0 else None)
    if outs[0] is not None and outs[1] is not None:
        print("features bit-identical:", bool(torch.equal(outs[0], outs[1])), "max abs diff", float((outs[0] - outs[1]).abs().max()))
