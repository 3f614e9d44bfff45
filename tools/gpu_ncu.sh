#!/bin/bash
# Full ncu captures; raw + source pages exported to CSV on the box, the (large) .ncu-rep files are NOT brought back.
mkdir -p gpurun_out /tmp/ncu
cap() { # name skip count cmd...
  name=$1; s=$2; c=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s $s -c $c -f -o /tmp/ncu/$name "$@" > gpurun_out/$name.log 2>&1
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  ncu -i /tmp/ncu/$name.ncu-rep --page details --csv > gpurun_out/$name.details.csv 2>/dev/null
}
cap vis_fwd 0 7 python tools/profile_vis.py 0 1
cap vis_bwd 43 12 python tools/profile_vis.py 0 1
cap resnet 53 12 python tools/profile_resnet.py 2
# per-instruction source page of ONE FF GEMM launch (W1 forward)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 5 -c 1 -f -o /tmp/ncu/one python tools/profile_vis.py 0 1 > /dev/null 2>&1
ncu -i /tmp/ncu/one.ncu-rep --page source --csv > gpurun_out/vis_w1.source.csv 2>/dev/null
ls -la gpurun_out /tmp/ncu
lscpu | grep -E "Model name|^CPU\(s\)" 
python -c "import numpy; numpy.show_config()" 2>&1 | grep -i -A2 "openblas configuration" | head -5
