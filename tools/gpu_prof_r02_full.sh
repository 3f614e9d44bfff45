mkdir -p gpurun_out /tmp/ncu
export SQ_KMEANS_GRAPH=0
full() { name=$1; k=$2; s=$3; shift 3
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$k" -s $s -c 1 -f -o /tmp/ncu/$name "$@" > gpurun_out/$name.log 2>&1
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  ncu -i /tmp/ncu/$name.ncu-rep --page details > gpurun_out/$name.details.txt 2>/dev/null
  wc -c gpurun_out/$name.details.txt; }
full r02_full_convgemm_L1c3 "convgemm_kernel" 55 python tools/profile_resnet.py 2      # layer-1 conv3 (64->256 + residual) of the second batch
full r02_full_convgemm_L1c2_halo "convgemm_kernel" 53 python tools/profile_resnet.py 2
full r02_full_convgemm_L3c2 "convgemm_kernel" 80 python tools/profile_resnet.py 2
full r02_full_vis_gemm_fuse3 "gemm_tc_kernel" 5 python tools/profile_vis.py 0 1
full r02_full_uni_gemm_fc1 "gemm_tc_kernel" 6 python tools/profile_uni.py 1
full r02_full_km_dist "km_dist_kernel" 5 python tools/profile_kmeans.py 1
