#!/bin/bash
timeout 400 python -m pytest tests/test_gemm_gpu.py tests/test_vis_gpu.py -m gpu -q -x 2>&1 | tail -3
for b in 1 0; do echo "SQ_BN192=$b"; SQ_BN192=$b timeout 200 python tools/profile_vis.py 5 3 2>&1 | tail -2; done
