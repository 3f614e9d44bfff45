#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
i=0
for cfg in "nccl 8" "multimem 8" "multimem 16" "multimem 4"; do
  set -- $cfg; i=$((i+1))
  timeout 100 $TR --master-port $((29560+i)) tools/dp_study.py --comm $1 --ctas $2 --steps 20 2>gpurun_out/dp8_$1_$2.err | tail -1 | tee gpurun_out/dp8_$1_$2.json | cut -c1-330
done
