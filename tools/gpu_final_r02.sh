#!/bin/bash
# Closing run of round 2 on one B200: full GPU test suite, smoke(), both bench arms, ncu evidence for the kernels added in this
# session (stem_tc_kernel, uni_attention_tc_kernel) and a fresh per-launch metric list of a ResNet batch and a UNI forward.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/ncu
timeout 400 python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r02_pytest_gpu_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tc.sum,sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
met() { name=$1; k=$2; s=$3; c=$4; shift 4
  timeout 400 ncu --metrics $M --clock-control none -k regex:"$k" -s $s -c $c --csv --log-file gpurun_out/$name.csv "$@" > gpurun_out/$name.log 2>&1
  python tools/ncu_table.py gpurun_out/$name.csv --json gpurun_out/$name.json > gpurun_out/$name.txt 2>&1
  tail -2 gpurun_out/$name.txt; }
full() { name=$1; k=$2; s=$3; shift 3
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$k" -s $s -c 1 -f -o /tmp/ncu/$name "$@" > gpurun_out/$name.log 2>&1
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  ncu -i /tmp/ncu/$name.ncu-rep --page details > gpurun_out/$name.details.txt 2>/dev/null
  grep -E "Duration|DRAM Throughput|Memory Throughput  " gpurun_out/$name.details.txt | head -3; }
met r02_resnet_batch64_metrics_final "stem|convgemm|bneck|avgpool|maxpool" 49 49 python tools/profile_resnet.py 2      # second batch: stem + 48 convolution launches (layer 1: conv2 + conv3 [+ downsample] are one kernel)
full r02_full_bneck_l1 "bneck_l1" 4 python tools/profile_resnet.py 2
full r02_full_bneck_l1_ds "bneck_l1" 3 python tools/profile_resnet.py 2
python tools/traffic_json.py gpurun_out/r02_resnet_batch64_metrics_final.json profiles/r02_traffic.json
cp profiles/r02_traffic.json gpurun_out/r02_traffic.json
timeout 600 python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; echo "bench rc=$?"; head -c 600 gpurun_out/r02_bench_n1_final.json; echo
timeout 400 python bench.py --impl reference > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "ref rc=$?"; head -c 400 gpurun_out/r02_bench_reference_arm.json; echo
