#!/bin/bash
# ncu --set full of the label-path kernels (surface mask/emit, voxel, ICP) from one short bench run
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'icp|voxel|surface' -c 10 -o gpurun_out/prof_label -f \
    python bench.py --steps 4 --warmup 1 --no-train > gpurun_out/ncu_label.log 2>&1
tail -3 gpurun_out/ncu_label.log | cut -c1-300
