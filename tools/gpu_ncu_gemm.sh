#!/bin/bash
# ncu captures: launch list of one bench run + full capture of one step's GEMM launches.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-icp > gpurun_out/ncu_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_split -s 36 -c 12 -o gpurun_out/prof_gemm -f \
    python bench.py --steps 2 --warmup 1 --no-icp > gpurun_out/ncu_gemm.log 2>&1
tail -3 gpurun_out/ncu_gemm.log
ls -la gpurun_out
