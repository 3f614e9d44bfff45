#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gemm_gpu.py -m gpu -x -q -s -k "block_diagonal or split_k or epilogues or timing_report" > gpurun_out/pytest_gemm.log 2>&1; grep -E "gemm timing|passed|failed|Error" gpurun_out/pytest_gemm.log
python -m pytest tests/test_vis_gpu.py -m gpu -q -x > gpurun_out/pytest_vis.log 2>&1; tail -5 gpurun_out/pytest_vis.log
for bn in 0 256; do echo "SQ_GEMM_BN=$bn"; SQ_GEMM_BN=$bn timeout 300 python tools/profile_vis.py 3 3 2>&1 | tail -2; done | tee gpurun_out/vis_timing.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_vis.csv python tools/profile_vis.py 1 1 > gpurun_out/ncu_vis.log 2>&1
