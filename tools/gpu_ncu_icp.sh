#!/bin/bash
# ncu --set full of the ICP and voxel-grid kernels from one short bench run
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'icp|voxel' -c 6 -o gpurun_out/prof_icp -f \
    python bench.py --steps 4 --warmup 1 --no-train > gpurun_out/ncu_icp.log 2>&1
tail -2 gpurun_out/ncu_icp.log | cut -c1-200
