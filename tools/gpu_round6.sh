#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json"))
print("resnet value", d["value"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], "share", d["roofline"]["kernel_share_of_step"])
v = d["vis_train"]; print("vis", v["value"], "ms", v["ms_per_step"], "e2e", v["e2e"]["value"], "issued frac", v["roofline"]["issued_frac_of_peak"], "share", v["roofline"]["kernel_share_of_step"])
print("kmeans", d["kmeans"]["value"], d["kmeans"]["ms_per_slide"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 109 -c 56 --csv --log-file gpurun_out/launches_resnet.csv python tools/profile_resnet.py 3 > gpurun_out/ncu_launch.log 2>&1
