#!/bin/bash
# compute-sanitizer memcheck over small configurations of every path (slow: keep the selection small)
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 "$@" > gpurun_out/sanitize_$name.log 2>&1; echo "$name rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitize_$name.log | tr '\n' ' ')"; }
run gemm python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -k "block_diagonal or epilogues or split_k or stream_k_scheduling or 192"
run vis python -m pytest tests/test_vis_gpu.py -m gpu -q -x -k "gradients_match_oracle_and_golden and small or ragged or accumulation"
run resnet python -m pytest tests/test_resnet_gpu.py -m gpu -q -x -k "golden"
run uni python -m pytest tests/test_uni_gpu.py -m gpu -q -x -k "extract_matches_oracle and 2-3"
run metrics python -m pytest tests/test_metrics_gpu.py tests/test_kmeans_gpu.py -m gpu -q -x -k "small or b2 or odd or edge"
