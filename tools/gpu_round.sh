#!/bin/bash
# Full GPU pass: parity tests, smoke, bench (both arms), ncu launch list + full captures.
mkdir -p gpurun_out
L=gpurun_out/round.log
nvidia-smi -L > $L 2>&1
echo "== pytest -m gpu" >> $L
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -30 >> $L
echo "== smoke" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
echo "== bench reference arm" >> $L
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> $L
echo "== bench" >> $L
timeout 900 python bench.py > gpurun_out/bench.json 2>> $L
cat gpurun_out/bench_ref.json gpurun_out/bench.json >> $L
if [ "$1" != "noprof" ]; then
echo "== ncu launch list" >> $L
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-icp > gpurun_out/ncu_bench.log 2>&1
echo "== ncu full: gemm" >> $L
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_split -s 36 -c 12 -o gpurun_out/prof_gemm -f \
    python bench.py --steps 2 --warmup 1 --no-icp > gpurun_out/ncu_gemm.log 2>&1
tail -3 gpurun_out/ncu_gemm.log >> $L
fi
tail -120 $L
