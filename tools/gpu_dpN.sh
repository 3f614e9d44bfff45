#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 1 --warmup 3 > gpurun_out/bench_ref_n$N.json 2>> gpurun_out/bench_n$N.err; echo "ref rc=$?"
python - <<PY
import json
txt = open("gpurun_out/bench_n$N.json").read()
print("stdout lines:", len([l for l in txt.splitlines() if l.strip()]))
d = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
print("N=$N resnet", round(d["value"]), "e2e", round(d["e2e"]["value"]), "| vis", round(d["vis_train"]["value"]), d["vis_train"]["ms_per_step"], "e2e", round(d["vis_train"]["e2e"]["value"]), "| kmeans", round(d["kmeans"]["value"]), "| uni", round(d["uni_extract"]["value"]))
r = open("gpurun_out/bench_ref_n$N.json").read()
print("ref stdout lines:", len([l for l in r.splitlines() if l.strip()]), json.loads(r.splitlines()[0])["value"])
PY
