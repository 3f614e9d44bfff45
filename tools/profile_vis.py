"""Timing / profiling driver for the ViS train step (BASELINE configs[2]: 32 slides, 100x2048 -> 20530 genes, AdamW)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vis_oracle as V  # noqa: E402  (weight / input generators only)
from sequoia_pub_b200.tformer_lin import ViS  # noqa: E402
from sequoia_pub_b200.train import FusedTrainer  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B, G = 32, 20530
torch.manual_seed(0)
m = ViS(num_outputs=G, input_dim=2048, depth=6, nheads=16, dimensions_f=64, dimensions_s=64, dimensions_c=64).cuda().train()
x, y = V.make_inputs(0, B, G)
x, y = x.cuda(), y.cuda()
tr = FusedTrainer(m, lr=1e-3, weight_decay=0.0)
for _ in range(warm):
    tr.step(x, y)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
s.record()
for _ in range(steps):
    loss = tr.step(x, y)
e.record()
t_cpu = time.perf_counter() - t0
torch.cuda.synchronize()
ms = s.elapsed_time(e) / max(steps, 1)
if steps: print(f"[vis train] {ms:.3f} ms/step (host enqueue {t_cpu / max(steps,1) * 1e3:.3f} ms/step) -> {B / ms * 1e3:.1f} slides/s, loss {loss.item():.5f}")
if steps:
    m.eval()
    with torch.no_grad():
        for _ in range(3):
            m(x)
        s.record()
        for _ in range(steps):
            m(x)
        e.record()
    torch.cuda.synchronize()
    print(f"[vis forward] {s.elapsed_time(e) / steps:.3f} ms/batch of {B}")
