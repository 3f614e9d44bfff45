#!/bin/bash
# One GPU-box visit: parity tests, bench, launch list, full ncu capture of the conv kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 109 -c 112 --csv --log-file gpurun_out/launches_resnet.csv python tools/profile_resnet.py 3 > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 53 -c 53 -o gpurun_out/prof_resnet_conv -f python tools/profile_resnet.py 3 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
