#!/bin/bash
# Per-launch ncu metrics of ONE batch-64 ResNet-50 forward (second batch: weights prepacked, caches cold per ncu replay).
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tc.sum,sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
timeout 900 ncu --metrics $M --clock-control none -k regex:"convgemm|gemm_tc|stem_fused|avgpool" -s 54 -c 54 --csv --log-file gpurun_out/resnet_metrics.csv python tools/profile_resnet.py 2 > gpurun_out/resnet_metrics.log 2>&1
python tools/ncu_table.py gpurun_out/resnet_metrics.csv --json gpurun_out/resnet_metrics.json | tee gpurun_out/resnet_metrics.txt | tail -60
