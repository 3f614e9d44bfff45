#!/bin/bash
# 8-GPU call: BASELINE configs[4] end to end (256 WSIs), the bench line at N=8 (headline + DP ViS step + configs[3] UNI leg) and the
# 2-GPU NCCL parity tests.  Everything under `timeout`.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $TR --nproc-per-node 8 --master-port 29541 tools/e2e_config5.py --slides 256 --train-steps 10 > gpurun_out/e2e_config5_n8.json 2> gpurun_out/e2e_config5_n8.err
tail -c 700 gpurun_out/e2e_config5_n8.json
timeout 240 $TR --nproc-per-node 8 --master-port 29542 bench.py --gpus 8 --steps 2 --warmup 3 --only vis,uni > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
tail -c 1500 gpurun_out/bench_n8.json
CUDA_VISIBLE_DEVICES=0,1 timeout 200 python -m pytest tests/test_dp_gpu.py -m gpu -q --timeout 150 2>&1 | tail -3
