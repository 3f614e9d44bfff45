#!/bin/bash
# 2-GPU closing call: NCCL / multimem data-parallel parity tests and the bench line at N=2 (headline + DP ViS step).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 250 python -m pytest tests/test_dp_gpu.py -m gpu -q --timeout 200 > gpurun_out/r02_pytest_dp2.log 2>&1; echo "dp2 rc=$?"; tail -2 gpurun_out/r02_pytest_dp2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 3 --warmup 3 --only vis > gpurun_out/r02_bench_n2_final.json 2> gpurun_out/r02_bench_n2_final.err; echo "bench rc=$?"; tail -c 1800 gpurun_out/r02_bench_n2_final.json
