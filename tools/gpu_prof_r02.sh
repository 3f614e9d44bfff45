#!/bin/bash
# Round-2 ncu evidence (row N1): per-launch metrics (duration, DRAM bytes, L2 bytes, L2->SM bytes, tensor-pipe activity, DRAM %) of
# every kernel of the four hot paths as they are NOW, plus `--set full` captures of one launch per distinct kernel.
mkdir -p gpurun_out /tmp/ncu
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tc.sum,sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
met() { # name skip count cmd...
  name=$1; s=$2; c=$3; shift 3
  timeout 600 ncu --metrics $M --clock-control none -s $s -c $c --csv --log-file gpurun_out/$name.csv "$@" > gpurun_out/$name.log 2>&1
  python tools/ncu_table.py gpurun_out/$name.csv --json gpurun_out/$name.json > gpurun_out/$name.txt 2>&1
  tail -2 gpurun_out/$name.txt
}
# SQ_KMEANS_GRAPH=0: ncu cannot replay kernels inside a conditional graph node; the host loop launches the same kernels
export SQ_KMEANS_GRAPH=0
met r02_resnet_batch64_metrics 108 54 python tools/profile_resnet.py 2            # second batch (after 53 prepack + 54 first-batch launches + 1 memset-free)
met r02_vis_train_step_metrics 0 400 python tools/profile_vis.py 0 2              # skip handled below (second step selected by the table tool)
met r02_uni_batch64_metrics 0 120 python tools/profile_uni.py 1
met r02_kmeans_metrics 0 260 python tools/profile_kmeans.py 1
met r02_vit_train_step_metrics 0 260 python tools/profile_vit.py
# --set full captures: one launch per distinct kernel (-k matches the base name; -s picks the launch) (raw page -> csv; the .ncu-rep files stay on the box)
full() { # name kernel-regex skip cmd...
  name=$1; k=$2; s=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$k" -s $s -c 1 -f -o /tmp/ncu/$name "$@" > gpurun_out/$name.log 2>&1
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  ncu -i /tmp/ncu/$name.ncu-rep --page details > gpurun_out/$name.details.txt 2>/dev/null
}
full r02_full_convgemm_L1c3 "convgemm_kernel" 55 python tools/profile_resnet.py 2      # layer-1 conv3 (64->256 + residual) of the second batch
full r02_full_convgemm_L1c2_halo "convgemm_kernel" 53 python tools/profile_resnet.py 2
full r02_full_convgemm_L3c2 "convgemm_kernel" 80 python tools/profile_resnet.py 2
full r02_full_stem "stem_fused" 1 python tools/profile_resnet.py 2
full r02_full_vis_gemm_fuse3 "gemm_tc_kernel" 5 python tools/profile_vis.py 0 1
full r02_full_adamw "adamw" 0 python tools/profile_vis.py 0 1
full r02_full_uni_attention "uni_attention" 1 python tools/profile_uni.py 1
full r02_full_uni_gemm_fc1 "gemm_tc_kernel" 6 python tools/profile_uni.py 1
full r02_full_km_select "km_select" 5 python tools/profile_kmeans.py 1
full r02_full_km_assign "km_assign" 1 python tools/profile_kmeans.py 1
full r02_full_km_dist "km_dist_kernel" 5 python tools/profile_kmeans.py 1
ls -la gpurun_out | grep r02_ | wc -l
