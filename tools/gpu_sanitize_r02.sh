#!/bin/bash
# compute-sanitizer memcheck of the kernels added at the end of round 2: tcgen05 stem (stem_tc_kernel), fused layer-1 bottleneck tail
# (bneck_l1_kernel), tcgen05 attention (uni_attention_tc_kernel).  Small configurations only (memcheck is 10-50x slower).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 "$@" > gpurun_out/sanitize_r02_$name.log 2>&1; echo "$name rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitize_r02_$name.log | tr '\n' ' ')"; }
run bneck python -m pytest tests/test_conv_gpu.py -m gpu -q -x -k "fused_layer1_tail and (1-16-8 or 3-32-24)"
run resnet python -m pytest tests/test_resnet_gpu.py -m gpu -q -x -k "extract_matches_golden"
run uni python -m pytest tests/test_uni_gpu.py -m gpu -q -x -k "extract_matches_oracle and 2-3"
cat gpurun_out/sanitize_r02_*.log | grep -E "ERROR SUMMARY|Invalid|out of bounds" | head -10 > gpurun_out/r02_compute_sanitizer_memcheck.txt; cat gpurun_out/r02_compute_sanitizer_memcheck.txt
