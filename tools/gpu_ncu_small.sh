#!/bin/bash
# ncu full capture of the non-GEMM kernels of one bench step (skip warm-up launches)
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'frontend|dense|pool|pose_|refiner_out|posenet_out' -s 40 -c 16 -o gpurun_out/prof_small -f \
    python bench.py --steps 2 --warmup 1 --no-icp > gpurun_out/ncu_small.log 2>&1
tail -3 gpurun_out/ncu_small.log
IMPLS=0 TAIL=25 bash tools/gpu_gemm_ab.sh
