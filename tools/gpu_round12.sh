#!/bin/bash
timeout 500 python -m pytest tests/test_gemm_gpu.py tests/test_resnet_gpu.py tests/test_vis_gpu.py -m gpu -q -x 2>&1 | tail -2
for v in 1 0; do echo "SQ_PDL=$v"; SQ_PDL=$v timeout 100 python tools/host_overhead.py 2>&1 | grep "C call"; SQ_PDL=$v timeout 200 python tools/profile_vis.py 5 3 2>&1 | tail -2; done
