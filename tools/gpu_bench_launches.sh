#!/bin/bash
# ncu launch list of the bench command itself (profiling recipe: --metrics gpu__time_duration.sum --clock-control none): the first 400
# kernel launches of `python bench.py` after the weight prepack, i.e. eight batches of stem + 49 convolution launches.  Per-launch
# times are cold-cache and serialised; what must agree with the bench line is each kernel's SHARE of the batch.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"stem|convgemm|bneck" -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv \
    python bench.py --steps 1 --warmup 3 --only none --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.json 2> gpurun_out/r02_bench_under_ncu.err
echo "rc=$?"
python tools/summarize_launches.py gpurun_out/r02_bench_launches.csv > gpurun_out/r02_bench_launches_summary.txt 2>&1; tail -12 gpurun_out/r02_bench_launches_summary.txt
