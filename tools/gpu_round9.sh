#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -k "stream_k or split or epilogues" 2>&1 | tail -5
timeout 400 python -m pytest tests/test_vis_gpu.py -m gpu -q -x 2>&1 | tail -3
for sk in 1 0; do echo "SQ_STREAMK=$sk"; SQ_STREAMK=$sk timeout 200 python tools/profile_vis.py 5 3 2>&1 | tail -2; done
