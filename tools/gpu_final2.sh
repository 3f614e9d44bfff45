#!/bin/bash
# Round-1 closing run after the fused stem / ViT / HDF5 work: full GPU parity suite, N=1 bench line, ResNet launch list.
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json"))
print("resnet value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "roof", round(d["roofline"]["frac"], 3), "share", round(d["roofline"]["kernel_share_of_step"], 2), d["clocks"])
v = d["vis_train"]; print("vis", round(v["value"]), "ms", round(v["ms_per_step"], 3), "e2e", round(v["e2e"]["value"]), "alg frac", round(v["roofline"]["frac"], 3))
print("kmeans", round(d["kmeans"]["value"], 1)); u = d["uni_extract"]; print("uni", round(u["value"]), "roof", round(u["roofline"]["frac"], 3))
t = d["vit_train"]; print("vit", round(t["value"]), "ms", round(t["ms_per_step"], 3))
print("cpu resnet", d["cpu_baseline"])
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 107 -c 54 --csv --log-file gpurun_out/launches_resnet.csv python tools/profile_resnet.py 3 > /dev/null 2>&1
tail -2 gpurun_out/launches_resnet.csv | cut -c1-200
