#!/bin/bash
# ncu evidence for one bf16 training step (cfg 2): launch list + full captures of the two pair kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --profile --no-graph --steps 1 --warmup 3 > gpurun_out/launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cfconv_pair_kernel -s 30 -c 1 -f -o gpurun_out/pair_fwd \
    python bench.py --profile --no-graph --steps 1 --warmup 3 > gpurun_out/prof1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cfconv_fused_bwd_kernel -s 15 -c 1 -f -o gpurun_out/pair_bwd \
    python bench.py --profile --no-graph --steps 1 --warmup 3 > gpurun_out/prof2.log 2>&1
ls -la gpurun_out/*.ncu-rep
