"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("ns", "nsecond"):
            v /= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            v *= 1e3
        rows.append((r["Kernel Name"].split("(")[0], v, r["Grid Size"], r["Block Size"]))
rows = rows[skip:]
agg = defaultdict(lambda: [0, 0.0])
for k, v, *_ in rows:
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v for _, v, *_ in rows)
print(f"{len(rows)} launches, {tot/1e3:.3f} ms total (serialised, cold-cache)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t/tot*100:6.2f}%  {t:10.1f} us  n={n:4d}  avg {t/n:8.1f} us  {k}")
if len(sys.argv) > 3:
    for i, (k, v, g, b) in enumerate(rows):
        print(i, f"{v:9.1f}", g, b, k)
