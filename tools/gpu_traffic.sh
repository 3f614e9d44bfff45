#!/bin/bash
# DRAM traffic of the dominant kernel (gemm_tc_kernel) per launch: ResNet batch-64 forward and one ViS train step.
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:gemm_tc -s 53 -c 53 --csv --log-file gpurun_out/traffic_resnet.csv python tools/profile_resnet.py 2 > /dev/null 2>&1
timeout 300 ncu --metrics $M --clock-control none -k regex:gemm_tc -s 129 -c 129 --csv --log-file gpurun_out/traffic_vis.csv python tools/profile_vis.py 0 2 > /dev/null 2>&1
python - <<'PY'
import csv, json
out = {}
for name in ("resnet", "vis"):
    rows = [r for r in csv.DictReader(l for l in open(f"gpurun_out/traffic_{name}.csv") if l.startswith('"'))]
    tot = {"dram__bytes_read.sum": 0.0, "dram__bytes_write.sum": 0.0, "gpu__time_duration.sum": 0.0}
    ids = set()
    for r in rows:
        v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(u, 1)
        tot[r["Metric Name"]] += v * mult; ids.add(r["ID"])
    n = len(ids)
    out[name] = {"launches": n, "dram_read_bytes": tot["dram__bytes_read.sum"], "dram_write_bytes": tot["dram__bytes_write.sum"],
                 "traffic_bytes_per_launch": (tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]) / max(n, 1), "sum_duration_us": tot["gpu__time_duration.sum"]}
json.dump(out, open("gpurun_out/traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
PY
