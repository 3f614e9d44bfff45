#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 200 python -m pytest tests/test_dp_gpu.py -m gpu -q -s 2>&1 | grep -E "dp2|passed|failed"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"
python - <<'PY'
import json
txt = open("gpurun_out/bench_n2.json").read()
d = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
print("lines on stdout:", len([l for l in txt.splitlines() if l.strip()]))
print("N=2 resnet value", round(d["value"]), "e2e", round(d["e2e"]["value"]))
v = d["vis_train"]; print("vis", round(v["value"]), "ms", v["ms_per_step"], "e2e", round(v["e2e"]["value"]))
print("kmeans", d["kmeans"]["value"], "uni", d["uni_extract"]["value"])
PY
