#!/bin/bash
# A/B of the tcgen05 stem (SQ_STEM_TC=1) against the mma.sync fused stem (SQ_STEM_TC=0): parity tests, then batch time with
# 1 / 2 / 3 extractor lanes for both settings on the same box.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SQ_STEM_TC=1 timeout 150 python -m pytest tests/test_resnet_gpu.py -m gpu -x -q -s > gpurun_out/stemtc_1.log 2>&1
echo "pytest SQ_STEM_TC=1 rc=$?"; grep -E "timing|passed|failed|rror|timeout" gpurun_out/stemtc_1.log | head -8
for rep in 1 2; do for v in 0 1; do echo "== SQ_STEM_TC=$v"; SQ_STEM_TC=$v timeout 100 python tools/lanes_compare.py 2>&1 | tail -3; done; done
