#!/bin/bash
# A/B of the stem kernels: SQ_STEM_TC=0 mma.sync fused stem, 1 tcgen05 stem (im2col pairs in tensor memory).
# Parity tests, per-role wait counters, then batch time with 1 / 2 lanes.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in ${STEM_VARIANTS:-1}; do
  SQ_STEM_TC=$v timeout 150 python -m pytest tests/test_resnet_gpu.py -m gpu -x -q -s > gpurun_out/stemtc_$v.log 2>&1
  echo "pytest SQ_STEM_TC=$v rc=$?"; grep -E "timing|passed|failed|rror|timeout" gpurun_out/stemtc_$v.log | head -8
  SQ_STEM_TC=$v timeout 100 python tools/stem_prof.py 2>&1 | tail -5
done
for rep in 1 2; do for v in 0 ${STEM_VARIANTS:-1}; do echo "== SQ_STEM_TC=$v"; SQ_STEM_TC=$v timeout 100 python tools/lanes_compare.py 2>&1 | tail -3 | head -2; done; done
