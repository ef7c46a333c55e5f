#!/bin/bash
# GEMM bring-up / tuning pass: parity of the network tests, then per-layer timings per implementation / tile-width mask.
mkdir -p gpurun_out
L=gpurun_out/gemm_ab.log
nvidia-smi -L > $L 2>&1
timeout 600 python -m pytest tests/test_gpu_net.py -x -q -m gpu 2>&1 | tail -15 >> $L
for impl in ${IMPLS:-0 3}; do
for mask in ${MASKS:-0x3f}; do
  echo "== impl $impl mask $mask" >> $L
  APE_GEMM_IMPL=$impl APE_GEMM_WIDE_MASK=$mask timeout 300 python bench.py --steps 50 --warmup 5 --no-icp > gpurun_out/bench_${impl}_$mask.json 2>> $L
  python tools/bench_layers.py gpurun_out/bench_${impl}_$mask.json >> $L 2>&1
done
done
tail -${TAIL:-90} $L
