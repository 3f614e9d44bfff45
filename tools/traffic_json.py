"""Rewrites the `resnet` entry of profiles/r02_traffic.json (DRAM bytes per convolution launch, what bench.py reports as
roofline.traffic) from a per-launch ncu metrics JSON (tools/ncu_table.py --json) of one ResNet batch."""
import json
import sys

src, dst = sys.argv[1], sys.argv[2]
recs = [r for r in json.load(open(src)) if "convgemm" in r["kernel"] or "bneck" in r["kernel"]]
total = sum(r["dram_rd_MB"] + r["dram_wr_MB"] for r in recs) * 1e6
d = json.load(open(dst))
d["resnet"] = {"launches": len(recs), "traffic_bytes_per_launch": total / len(recs), "dram_MB_per_batch64": total / 1e6, "dram_MB_per_patch": total / 1e6 / 64,
               "source": f"{src} (ncu --metrics, second batch of tools/profile_resnet.py 2: convgemm_kernel + bneck_l1_kernel launches)"}
json.dump(d, open(dst, "w"), indent=1)
print(f"resnet: {len(recs)} launches, {total / 1e6:.1f} MB per batch of 64, {total / len(recs) / 1e6:.2f} MB per launch")
