#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_resnet_gpu.py tests/test_gemm_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 109 -c 56 --csv --log-file gpurun_out/launches_resnet.csv python tools/profile_resnet.py 3 > gpurun_out/ncu_launch.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_resnet.csv 0 x | head -40
timeout 200 python tools/host_overhead.py
