import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import resnet50_oracle as O
from sequoia_pub_b200.resnet import resnet50
from sequoia_pub_b200 import _lib
import ctypes as C
m = resnet50().eval(); m.load_state_dict(O.make_state_dict(0)); m = m.cuda()
x = torch.randint(0, 256, (64, 256, 256, 3), dtype=torch.uint8, device="cuda")
out = torch.empty(64, 2048, device="cuda")
for _ in range(3): m.extract_uint8(x, out=out)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): m.extract_uint8(x, out=out)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"extract_uint8: host enqueue {(t1-t0)/20*1e3:.3f} ms/batch, total {(t2-t0)/20*1e3:.3f} ms/batch")
# pieces
t0 = time.perf_counter()
for _ in range(20): m._prepack()
print(f"_prepack check: {(time.perf_counter()-t0)/20*1e3:.3f} ms")
L = _lib.lib(); pw, sh = m._prepack(); ws = m._workspace
t0 = time.perf_counter()
for _ in range(20):
    _lib.check(L.sq_resnet50_extract(_lib.ptr(x), 0, 64, 256, 256, _lib.ptr(pw), _lib.ptr(sh), _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"C call only: host enqueue {(t1-t0)/20*1e3:.3f} ms/batch, total {(t2-t0)/20*1e3:.3f} ms/batch")
