import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import resnet50_oracle as O
from sequoia_pub_b200.resnet import resnet50
m = resnet50().eval(); m.load_state_dict(O.make_state_dict(0)); m = m.cuda()
x = torch.randint(0, 256, (1024, 256, 256, 3), dtype=torch.uint8, device="cuda")
ref = torch.cat([m.extract_uint8(x[b:b + 64]) for b in range(0, 1024, 64)])
for lanes in (1, 2, 3):
    out = m.extract_many(x, lanes=lanes)
    torch.cuda.synchronize()
    assert torch.equal(out, ref), lanes
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(3): m.extract_many(x, out=out, lanes=lanes)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 3
    print(f"lanes={lanes}: {ms / 16:.3f} ms/batch -> {1024 / ms * 1e3:.0f} patches/s")
