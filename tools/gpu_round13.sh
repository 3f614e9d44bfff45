#!/bin/bash
# Fused stem (word-load staging) parity + timing, then ncu --set full captures of the new kernels (stem, ViT attention).
mkdir -p gpurun_out /tmp/ncu
timeout 200 python -m pytest tests/test_resnet_gpu.py -m gpu -x -q -s 2>&1 | tail -12 | tee gpurun_out/pytest_stem.log
cap() { # name kernel-regex skip count cmd...
  name=$1; k=$2; s=$3; c=$4; shift 4
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -f -o /tmp/ncu/$name "$@" > gpurun_out/$name.log 2>&1
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  ncu -i /tmp/ncu/$name.ncu-rep --page details --csv > gpurun_out/$name.details.csv 2>/dev/null
}
cap stem stem_fused 1 1 python tools/profile_resnet.py 2
cap vit_attn vit_attn 5 2 python tools/profile_vit.py 1 1000
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_vit.csv python tools/profile_vit.py 1 > /dev/null 2>&1
ls -la gpurun_out | tail -8
