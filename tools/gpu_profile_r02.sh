#!/bin/bash
# Round-2 evidence: bench lines (both arms), ncu launch list of the headline step, ncu --set full captures of the GEMM /
# dense kernels of one step, of the ADD-S kernel, of the multi-label back-projection and of the ICP kernel.
# Reduce with:  python tools/summarize_profiles.py r02 --launches gpurun_out/r02_launches_step.csv --rep gemm=gpurun_out/r02_prof_gemm.ncu-rep \
#       --rep adds=gpurun_out/r02_prof_adds.ncu-rep --rep label=gpurun_out/r02_prof_label.ncu-rep --rep icp=gpurun_out/r02_prof_icp.ncu-rep --traffic
mkdir -p gpurun_out
L=gpurun_out/r02_profile.log
nvidia-smi -L > $L 2>&1
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2>> $L
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2>> $L
timeout 900 python bench.py > gpurun_out/r02_bench_200steps.json 2>> $L
echo "== ncu launch list (headline step)" >> $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_step.csv \
    python bench.py --steps 2 --warmup 1 --no-icp --no-train > gpurun_out/r02_ncu_a.log 2>&1
echo "== ncu full: gemm + dense" >> $L
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'gemm_split|dense_swapped' -s 51 -c 17 -o gpurun_out/r02_prof_gemm -f \
    python bench.py --steps 2 --warmup 1 --no-icp --no-train > gpurun_out/r02_ncu_b.log 2>&1
tail -2 gpurun_out/r02_ncu_b.log >> $L
echo "== ncu full: ADD-S" >> $L
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'add_metric|knn3_top1' -s 1 -c 4 -o gpurun_out/r02_prof_adds -f \
    python bench.py --steps 4 --warmup 1 --no-train --no-c4 > gpurun_out/r02_ncu_c.log 2>&1
tail -2 gpurun_out/r02_ncu_c.log >> $L
echo "== ncu full: back-projection (single- and multi-label)" >> $L
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'surface_|view_offsets' -s 12 -c 14 -o gpurun_out/r02_prof_label -f \
    python bench.py --steps 4 --warmup 1 --no-train --no-c4 --no-adds > gpurun_out/r02_ncu_d.log 2>&1
tail -2 gpurun_out/r02_ncu_d.log >> $L
echo "== ncu full: ICP / voxel grid" >> $L
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'icp_p2p|voxel_down' -c 4 -o gpurun_out/r02_prof_icp -f \
    python bench.py --steps 4 --warmup 1 --no-train --no-c4 --no-adds > gpurun_out/r02_ncu_e.log 2>&1
tail -2 gpurun_out/r02_ncu_e.log >> $L
# reduce on the box (the .ncu-rep files together exceed what gpurun brings back) and drop the reports
mkdir -p gpurun_out/profiles_r02
python tools/summarize_profiles.py r02 --outdir gpurun_out/profiles_r02 --launches gpurun_out/r02_launches_step.csv \
    --rep gemm=gpurun_out/r02_prof_gemm.ncu-rep --rep adds=gpurun_out/r02_prof_adds.ncu-rep \
    --rep label=gpurun_out/r02_prof_label.ncu-rep --rep icp=gpurun_out/r02_prof_icp.ncu-rep >> $L 2>&1
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out gpurun_out/profiles_r02 >> $L
tail -30 $L
