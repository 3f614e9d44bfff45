#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gemm_gpu.py tests/test_vis_gpu.py -m gpu -q -x 2>&1 | tail -4
for cfg in "1 1" "1 0"; do set -- $cfg; echo "SQ_FUSE3=$1 SQ_STREAMK=$2"; SQ_FUSE3=$1 SQ_STREAMK=$2 timeout 200 python tools/profile_vis.py 5 3 2>&1 | tail -2; done

