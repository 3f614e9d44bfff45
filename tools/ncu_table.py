"""Turns an `ncu --csv --page raw` (or --metrics ... --csv) log into one line per launch: kernel, grid, duration, DRAM bytes,
L2 bytes, tensor-pipe activity.  Usage: python tools/ncu_table.py <csv> [--json out.json]"""
import csv
import json
import sys
from collections import defaultdict

WANT = {
    "gpu__time_duration.sum": "us",
    "dram__bytes_read.sum": "dram_rd_MB",
    "dram__bytes_write.sum": "dram_wr_MB",
    "lts__t_bytes.sum": "l2_MB",
    "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_sm_MB",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "hmma_pct",
    "sm__inst_executed_pipe_tc.sum": "tc_inst",
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum": "utchmma_ops",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
}
UNIT = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def parse(path):
    lines = [l for l in open(path, errors="replace") if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    out = []
    if rows and "Metric Name" in rows[0]:            # long format (one row per metric)
        by = defaultdict(dict)
        for r in rows:
            by[int(r["ID"])].update({"kernel": r["Kernel Name"].split("(")[0], "grid": r["Grid Size"], r["Metric Name"]: (r["Metric Value"], r["Metric Unit"])})
        recs = [by[k] for k in sorted(by)]
    else:                                            # wide format: second row holds the units
        units = rows[0]
        recs = []
        for r in rows[1:]:
            d = {"kernel": r["Kernel Name"].split("(")[0], "grid": r["Grid Size"]}
            for m in WANT:
                if m in r:
                    d[m] = (r[m], units.get(m, ""))
            recs.append(d)
    for d in recs:
        o = {"kernel": d["kernel"], "grid": d["grid"]}
        for m, name in WANT.items():
            if m in d:
                v, u = d[m]
                try:
                    x = float(str(v).replace(",", ""))
                except ValueError:
                    continue
                o[name] = x * UNIT.get(u, 1.0)
        out.append(o)
    return out


if __name__ == "__main__":
    recs = parse(sys.argv[1])
    cols = ["us", "dram_rd_MB", "dram_wr_MB", "l2_MB", "l2_to_sm_MB", "tensor_pct", "hmma_pct", "dram_pct", "sm_pct", "utchmma_ops", "tc_inst"]
    cols = [c for c in cols if any(c in r for r in recs)]
    print(f"{'#':>3} {'kernel':46s} {'grid':>12s} " + " ".join(f"{c:>11s}" for c in cols))
    for i, r in enumerate(recs):
        print(f"{i:3d} {r['kernel'][-46:]:46s} {r['grid']:>12s} " + " ".join(f"{r.get(c, float('nan')):11.4g}" for c in cols))
    tot = sum(r.get("us", 0) for r in recs)
    print(f"total {tot:.1f} us over {len(recs)} launches; DRAM {sum(r.get('dram_rd_MB', 0) + r.get('dram_wr_MB', 0) for r in recs):.0f} MB")
    if "--json" in sys.argv:
        json.dump(recs, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=0)
