#!/bin/bash
# First GPU call of round 2: validate and time the prepared EPI_CONV_PF epilogue against the default, then (if it is
# bit-identical and faster) the ResNet parity tests and the bench line with it switched on.
mkdir -p gpurun_out
timeout 300 python tools/conv_pf_check.py 2>&1 | tee gpurun_out/conv_pf_check.log
SQ_CONV_EPI_PF=1 timeout 200 python -m pytest tests/test_resnet_gpu.py tests/test_gemm_gpu.py -m gpu -q 2>&1 | tail -3 | tee -a gpurun_out/conv_pf_check.log
SQ_CONV_EPI_PF=1 timeout 300 python bench.py --steps 5 --warmup 3 --only none > gpurun_out/bench_pf.json 2> gpurun_out/bench_pf.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_pf.json"))
print("with EPI_CONV_PF: resnet value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "roof", round(d["roofline"]["frac"], 3))
PY
