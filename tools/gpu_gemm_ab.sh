#!/bin/bash
# GEMM bring-up / tuning pass: parity of the network tests, then per-layer timings for several tile-width masks.
mkdir -p gpurun_out
L=gpurun_out/gemm_ab.log
nvidia-smi -L > $L 2>&1
timeout 600 python -m pytest tests/test_gpu_net.py -x -q -m gpu 2>&1 | tail -15 >> $L
for mask in ${MASKS:-0x3f 0x00}; do
  echo "== mask $mask" >> $L
  APE_GEMM_WIDE_MASK=$mask timeout 300 python bench.py --steps 50 --warmup 5 --no-icp > gpurun_out/bench_$mask.json 2>> $L
  python tools/bench_layers.py gpurun_out/bench_$mask.json >> $L 2>&1
done
tail -80 $L
