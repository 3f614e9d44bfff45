#!/bin/bash
# Last run of round 1: host pipeline with 4 device tile buffers (bit-identity test) and the N=1 bench line of the final tree.
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_resnet_gpu.py tests/test_pipeline_gpu.py -m gpu -q -k "two_lane or file_level" 2>&1 | tail -2 | tee gpurun_out/pytest_lanes.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json"))
print("resnet value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "roof", round(d["roofline"]["frac"], 3), d["clocks"])
v = d["vis_train"]; print("vis", round(v["value"]), "e2e", round(v["e2e"]["value"])); print("kmeans", round(d["kmeans"]["value"], 1), "uni", round(d["uni_extract"]["value"]), "vit", round(d["vit_train"]["value"]))
PY
