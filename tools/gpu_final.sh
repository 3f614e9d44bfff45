#!/bin/bash
# Round-end style run: parity tests, smoke, both bench arms, launch lists (kept small enough to travel back).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json"))
print("resnet value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "roof", round(d["roofline"]["frac"], 3), "share", round(d["roofline"]["kernel_share_of_step"], 2), d["clocks"])
v = d["vis_train"]; print("vis", round(v["value"]), "ms", round(v["ms_per_step"], 3), "e2e", round(v["e2e"]["value"]), "alg frac", round(v["roofline"]["frac"], 3), "issued frac", round(v["roofline"]["issued_frac_of_peak"], 3), "cpu", v.get("cpu_baseline", {}).get("value"))
print("kmeans", round(d["kmeans"]["value"], 1), d["kmeans"]["ms_per_slide"], "cpu", d["kmeans"].get("cpu_baseline", {}).get("value"))
u = d["uni_extract"]; print("uni", round(u["value"]), "e2e", round(u["e2e"]["value"]), "roof", round(u["roofline"]["frac"], 3))
print("cpu resnet", d["cpu_baseline"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 109 -c 56 --csv --log-file gpurun_out/launches_resnet.csv python tools/profile_resnet.py 3 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_vis.csv python tools/profile_vis.py 0 2 > /dev/null 2>&1
