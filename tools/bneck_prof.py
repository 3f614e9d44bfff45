"""Per-role wait counters of the fused layer-1 bottleneck tail (SQ_BNECK_PROF=1), one batch of 64."""
import os, sys, torch
os.environ["SQ_BNECK_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import resnet50_oracle as O
from sequoia_pub_b200.resnet import resnet50
m = resnet50().eval(); m.load_state_dict(O.make_state_dict(0)); m = m.cuda()
x = torch.randint(0, 256, (128, 256, 256, 3), dtype=torch.uint8, device="cuda")
for r in range(2):
    m.extract_uint8(x[r * 64:(r + 1) * 64]); torch.cuda.synchronize(); print("--")
