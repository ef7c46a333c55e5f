#!/bin/bash
# Round-2 evidence after the one-launch refiner tail, the overlapped training all-reduce and the device-resident reconstruction
# loop: bench lines (both arms), ncu launch list of the headline step, ncu --set full of the step's GEMM / dense / tail launches
# and of the single-registration ICP kernel.  Summaries land in gpurun_out/profiles_r02d (copy into profiles/).
mkdir -p gpurun_out gpurun_out/profiles_r02d
L=gpurun_out/r02d_profile.log
nvidia-smi -L > $L 2>&1
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/profiles_r02d/r02d_bench_reference.json 2>> $L
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/profiles_r02d/r02d_bench.json 2>> $L
echo "== ncu launch list (headline step)" >> $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02d_launches_step.csv \
    python bench.py --steps 2 --warmup 1 --no-icp --no-train --no-c4 --no-adds > gpurun_out/r02d_ncu_a.log 2>&1
echo "== ncu full: gemm + dense + tail of one step" >> $L
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'gemm_split|dense_swapped|refiner_tail' -s 45 -c 15 -o gpurun_out/r02d_prof_gemm -f \
    python bench.py --steps 2 --warmup 1 --no-icp --no-train --no-c4 --no-adds > gpurun_out/r02d_ncu_b.log 2>&1
tail -2 gpurun_out/r02d_ncu_b.log >> $L
echo "== ncu full: single-registration ICP (512-thread CTA) in the reconstruction loop" >> $L
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'icp_p2p' -s 40 -c 3 -o gpurun_out/r02d_prof_icp1 -f \
    python tools/recon_profile.py > gpurun_out/r02d_ncu_c.log 2>&1
tail -2 gpurun_out/r02d_ncu_c.log >> $L
python tools/summarize_profiles.py r02d --outdir gpurun_out/profiles_r02d --launches gpurun_out/r02d_launches_step.csv \
    --rep gemm=gpurun_out/r02d_prof_gemm.ncu-rep --rep icp1=gpurun_out/r02d_prof_icp1.ncu-rep >> $L 2>&1
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out/profiles_r02d >> $L
tail -25 $L
