"""Maps the ncu launch list of one batch-64 ResNet-50 forward (profiles/r01_resnet_batch64_launches.txt, produced by
tools/summarize_launches.py ... list) onto the 52 bottleneck convolutions and prints, per launch, its FLOPs, the minimum
HBM bytes (input + output [+ residual] + weights, bf16), the larger of the two floors (bf16 sustained peak / HBM peak
from MEASURED_PEAKS.json) and how far the measured duration is from it."""
import sys

PEAK_TF, PEAK_GBS = 1364.7, 6545.0
path = sys.argv[1] if len(sys.argv) > 1 else "profiles/r01_resnet_batch64_launches.txt"
planes, blocks, strides = [64, 128, 256, 512], [3, 4, 6, 3], [1, 2, 2, 2]
order, inpl, H = [], 64, 64
for st in range(4):
    for b in range(blocks[st]):
        s = strides[st] if b == 0 else 1
        Ho = H // s
        seq = [("c1", (inpl, planes[st], 1, H, H)), ("c2", (planes[st], planes[st], 3, H, Ho))]
        if b == 0:
            seq.append(("ds", (inpl, planes[st] * 4, 1, H, Ho)))
        seq.append(("c3", (planes[st], planes[st] * 4, 1, Ho, Ho)))
        order += [(st + 1, b, n, c) for n, c in seq]
        inpl, H = planes[st] * 4, Ho
lines = [l.split() for l in open(path) if l[:1].isdigit() and "gemm_tc" in l]
assert len(lines) == len(order), (len(lines), len(order))
tot = ideal_tot = 0.0
print("conv         cin  cout k  Hin->Hout       M     GF  meas_us  TF/s  minMB  floor_us  x_floor  tile")
for (st, b, n, (cin, cout, k, Hi, Ho)), l in zip(order, lines):
    t = float(l[1])
    M, K = 64 * Ho * Ho, cin * k * k
    gf = 2 * M * cout * K / 1e9
    mb = (64 * Hi * Hi * cin * 2 + M * cout * 2 * (2 if n == "c3" else 1) + cout * K * 2) / 1e6
    floor = max(mb / PEAK_GBS * 1e3, gf / PEAK_TF * 1e3)
    tot += t
    ideal_tot += floor
    print(f"L{st}.b{b}.{n:3s} {cin:5d} {cout:5d} {k}  {Hi:3d}->{Ho:3d} {M:8d} {gf:6.1f} {t:8.1f} {gf / t * 1e3:5.0f} {mb:6.1f} {floor:9.1f} {t / floor:8.1f}  {' '.join(l[6:])[-22:]}")
print(f"sum of measured {tot:.0f} us, sum of floors {ideal_tot:.0f} us, ratio {tot / ideal_tot:.2f}")
