#!/bin/bash
# End-of-round evidence: bench lines (both arms), ncu launch lists (whole bench command, and the headline step alone),
# ncu --set full captures of the tcgen05 GEMM launches of one step, of the back-projection kernels and of the ICP /
# voxel kernels.  Everything lands in gpurun_out/; tools/summarize_profiles.py turns it into profiles/<tag>_*.
#   python tools/summarize_profiles.py r01f --launches gpurun_out/launches_step.csv --rep gemm=gpurun_out/prof_gemm.ncu-rep \
#          --rep label=gpurun_out/prof_label.ncu-rep --rep icp=gpurun_out/prof_icp.ncu-rep --traffic
mkdir -p gpurun_out
L=gpurun_out/profile_round.log
nvidia-smi -L > $L 2>&1
timeout 600 python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/bench_ref.json 2>> $L
timeout 900 python bench.py > gpurun_out/bench.json 2>> $L
echo "== ncu launch lists" >> $L
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_step.csv \
    python bench.py --steps 2 --warmup 1 --no-icp --no-train > gpurun_out/ncu_bench2.log 2>&1
echo "== ncu full: gemm" >> $L
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_split -s 36 -c 12 -o gpurun_out/prof_gemm -f \
    python bench.py --steps 2 --warmup 1 --no-icp --no-train > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/ncu_gemm.log >> $L
echo "== ncu full: back-projection" >> $L
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'surface_' -s 12 -c 9 -o gpurun_out/prof_label -f \
    python bench.py --steps 4 --warmup 1 --no-train > gpurun_out/ncu_label.log 2>&1
tail -2 gpurun_out/ncu_label.log >> $L
echo "== ncu full: ICP / voxel grid" >> $L
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'icp_p2p|voxel' -c 4 -o gpurun_out/prof_icp -f \
    python bench.py --steps 4 --warmup 1 --no-train > gpurun_out/ncu_icp.log 2>&1
tail -2 gpurun_out/ncu_icp.log >> $L
ls -la gpurun_out >> $L
tail -30 $L
