"""Profiling driver (run under ncu on the GPU box): a few batch-64 launches of the ResNet-50 extractor."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import resnet50_oracle as O  # noqa: E402  (weight generator only)
from sequoia_pub_b200.resnet import resnet50  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
m = resnet50().eval()
m.load_state_dict(O.make_state_dict(0))
m = m.cuda()
x = torch.randint(0, 256, (64 * reps, 256, 256, 3), dtype=torch.uint8, device="cuda")
for r in range(reps):
    m.extract_uint8(x[r * 64:(r + 1) * 64])
torch.cuda.synchronize()
