#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gemm_gpu.py -m gpu -x -q -k "block_diagonal" > gpurun_out/pytest_gemm.log 2>&1; tail -15 gpurun_out/pytest_gemm.log
python -m pytest tests/test_vis_gpu.py -m gpu -q -s > gpurun_out/pytest_vis.log 2>&1; tail -60 gpurun_out/pytest_vis.log
timeout 300 python tools/profile_vis.py 10 3 2>&1 | tee gpurun_out/vis_timing.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_vis.csv python tools/profile_vis.py 0 2 > gpurun_out/ncu_vis.log 2>&1
tail -2 gpurun_out/ncu_vis.log
