#!/bin/bash
# 8-GPU closing call: the bench line at N=8 (headline + DP ViS step with the multimem all-reduce).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 3 --warmup 3 --only vis --no-cpu-baseline > gpurun_out/r02_bench_n8_final.json 2> gpurun_out/r02_bench_n8_final.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r02_bench_n8_final.json
