#!/bin/bash
# Full ncu captures (small kernel counts so the reports stay < 64 MiB); raw pages exported to CSV on the box.
mkdir -p gpurun_out
cap() { # name skip count cmd...
  name=$1; s=$2; c=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s $s -c $c -f -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  ls -la gpurun_out/$name.ncu-rep
}
cap vis_fwd 0 7 python tools/profile_vis.py 0 1
cap vis_bwd 43 12 python tools/profile_vis.py 0 1
cap resnet 53 12 python tools/profile_resnet.py 2
