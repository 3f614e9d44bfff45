"""Profiling driver (run under ncu on the GPU box): train steps of the ViT softmax-attention baseline at the config-3 shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vis_oracle as V  # noqa: E402  (input generator only)
from sequoia_pub_b200.train import FusedTrainer  # noqa: E402
from sequoia_pub_b200.vit import ViT  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
G = int(sys.argv[2]) if len(sys.argv) > 2 else 20530
torch.manual_seed(0)
m = ViT(num_outputs=G, dim=2048, depth=6, heads=16, mlp_dim=2048, dim_head=64).cuda().train()
x, y = V.make_inputs(0, 32, G)
tr = FusedTrainer(m, lr=1e-3)
for _ in range(steps):
    tr.step(x.cuda(), y.cuda())
torch.cuda.synchronize()
