"""Maps the per-launch ncu metrics of one batch-64 ResNet-50 forward (the JSON written by tools/ncu_table.py --json, e.g.
profiles/r02_resnet_batch64_metrics_final.json: stem + 52 convolutions) onto the 52 bottleneck convolutions and prints, per
launch, its FLOPs, the minimum HBM bytes (input + output [+ residual] + weights, bf16), the larger of the two floors (bf16
sustained peak / HBM peak from MEASURED_PEAKS.json), how far the measured duration is from it, and what ncu saw (DRAM bytes,
L2 -> SM bytes, tensor-pipe activity)."""
import json
import sys

PEAK_TF, PEAK_GBS = 1364.7, 6545.0
path = sys.argv[1] if len(sys.argv) > 1 else "profiles/r02_resnet_batch64_metrics_final.json"
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
recs = json.load(open(path))
stem = [r for r in recs if "stem" in r["kernel"]]
convs = [r for r in recs if "convgemm" in r["kernel"] or "gemm_tc" in r["kernel"] or "bneck" in r["kernel"]]
if any("bneck" in r["kernel"] for r in convs):
    # layer 1's conv2 + conv3 run as one kernel (csrc/fusedconv.cuh): one table row "c23" per block; in the first block the downsample
    # runs inside it as well ("c23d", no residual tensor) when the list has one launch less
    with_ds = len(convs) == len(order) - 4
    merged = []
    for e in order:
        st, b, n, c = e
        if st == 1 and (n == "c2" or (with_ds and n == "ds")):
            continue
        merged.append((st, b, "c23d" if with_ds and b == 0 else "c23", c) if st == 1 and n == "c3" else e)
    order = merged
assert len(convs) == len(order), (len(convs), len(order))
tot = ideal_tot = 0.0
print("conv         cin  cout k  Hin->Hout       M     GF  meas_us  TF/s  DRAM_MB  L2>SM_MB  minMB  floor_us  x_floor  tensor%  kernel")
for (st, b, n, (cin, cout, k, Hi, Ho)), r in zip(order, convs):
    t = r["us"]
    M, K = 64 * Ho * Ho, cin * k * k
    gf = 2 * M * cout * K / 1e9
    mb = (64 * Hi * Hi * cin * 2 + M * cout * 2 * (2 if n == "c3" else 1) + cout * K * 2) / 1e6
    if n == "c23":                       # 3x3 64 -> 64 then 1x1 64 -> 256 + residual: the 64-channel intermediate never reaches HBM
        gf = (2 * M * 64 * 576 + 2 * M * 256 * 64) / 1e9
        mb = (M * 64 * 2 + 2 * M * 256 * 2 + (64 * 576 + 256 * 64) * 2) / 1e6
    if n == "c23d":                      # first block: + downsample(x) inside the kernel, no residual read
        gf = (2 * M * 64 * 576 + 2 * 2 * M * 256 * 64) / 1e9
        mb = (2 * M * 64 * 2 + M * 256 * 2 + (64 * 576 + 2 * 256 * 64) * 2) / 1e6
    floor = max(mb / PEAK_GBS * 1e3, gf / PEAK_TF * 1e3)
    tot += t
    ideal_tot += floor
    print(f"L{st}.b{b}.{n:4s}{cin:5d} {cout:5d} {k}  {Hi:3d}->{Ho:3d} {M:8d} {gf:6.1f} {t:8.1f} {gf / t * 1e3:5.0f} {r['dram_rd_MB'] + r['dram_wr_MB']:8.1f} "
          f"{r['l2_to_sm_MB']:9.1f} {mb:6.1f} {floor:9.1f} {t / floor:8.1f} {r['tensor_pct']:8.1f}  {r['kernel'].replace('void ', '')}")
print(f"sum of measured {tot:.0f} us, sum of floors {ideal_tot:.0f} us, ratio {tot / ideal_tot:.2f} (ncu durations are cold-cache and serialised: "
      f"the un-profiled chain overlaps through programmatic dependent launch)")
for r in stem:
    print(f"stem: {r['kernel']} {r['us']:.1f} us, DRAM read {r['dram_rd_MB']:.1f} MB, tensor pipe {r['tensor_pct']:.1f} %")
