mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_step.csv \
    python bench.py --steps 2 --warmup 1 --no-icp --no-train > gpurun_out/ncu_bench2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'icp_p2p|voxel' -c 4 -o gpurun_out/prof_icp -f \
    python bench.py --steps 4 --warmup 1 --no-train > gpurun_out/ncu_icp.log 2>&1
tail -2 gpurun_out/ncu_icp.log | cut -c1-200
