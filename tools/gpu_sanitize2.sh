#!/bin/bash
# compute-sanitizer memcheck over the kernels added late in round 1: fused ResNet stem (both staging paths), ViT attention
# forward/backward (small and reference-shaped case), file-level pipeline.
mkdir -p gpurun_out
run() { name=$1; shift; timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 "$@" > gpurun_out/sanitize_$name.log 2>&1; echo "$name rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitize_$name.log | tr '\n' ' ')"; }
run stem python -m pytest tests/test_resnet_gpu.py -m gpu -q -x -k "golden or batch_sizes"
run vit python -m pytest tests/test_vit_gpu.py -m gpu -q -x -k "forward_and_gradients"
