#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_dp_gpu.py -m gpu -q --timeout 140 -k multimem 2>&1 | tail -2
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
timeout 120 $TR --master-port 29553 tools/dp_study.py --comm multimem --ctas 16 --steps 10 2>gpurun_out/dp2_mm16.err | tail -1 | tee gpurun_out/dp2_mm16.json | cut -c1-400
timeout 120 $TR --master-port 29552 tools/dp_study.py --comm multimem --ctas 8 --steps 10 2>gpurun_out/dp2_mm.err | tail -1 | tee gpurun_out/dp2_mm.json | cut -c1-600
tail -3 gpurun_out/dp2_mm.err | cut -c1-300
