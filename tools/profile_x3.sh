#!/bin/bash
# ncu evidence for the fp32-grade fused mode (cfg 2): launch list of one step + full captures of the two x3 kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fp32.csv \
    python bench.py --precision fp32 --profile --no-graph --steps 1 --warmup 3 > gpurun_out/launches_fp32.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cfconv_dense_x3_kernel -s 14 -c 1 -f -o gpurun_out/x3_fwd \
    python bench.py --precision fp32 --profile --no-graph --steps 1 --warmup 3 > gpurun_out/prof_x3_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cfconv_dense_bwd_x3_kernel -s 8 -c 1 -f -o gpurun_out/x3_bwd \
    python bench.py --precision fp32 --profile --no-graph --steps 1 --warmup 3 > gpurun_out/prof_x3_bwd.log 2>&1
ls -la gpurun_out/*.ncu-rep
