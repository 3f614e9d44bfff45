import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import uni_oracle as U
from sequoia_pub_b200.uni import VisionTransformer
m = VisionTransformer().eval(); m.load_state_dict(U.make_state_dict(0)); m = m.cuda()
x = torch.randint(0, 256, (64, 224, 224, 3), dtype=torch.uint8, device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2): m.extract_uint8(x)
torch.cuda.synchronize()
