#!/bin/bash
# usage: gpu_ncu_one.sh <kernel regex> <out name> [launch skip]
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${3:-0} -c 1 -o gpurun_out/prof_$2 -f \
    python bench.py --steps 4 --warmup 1 > gpurun_out/ncu_$2.log 2>&1
tail -2 gpurun_out/ncu_$2.log | cut -c1-200
