#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L

timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"; tail -3 gpurun_out/bench_n2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_n2.json"))
print("N=2 resnet value", d["value"], "e2e", d["e2e"]["value"])
v = d["vis_train"]; print("vis", v["value"], "ms", v["ms_per_step"], "e2e", v["e2e"]["value"])
print("kmeans", d["kmeans"]["value"])
PY
