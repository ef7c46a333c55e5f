#!/bin/bash
# first GPU bring-up: SIMT-only kernels first, tcgen05 under a timeout
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/first.log 2>&1
for t in tests/test_gpu_backproject.py tests/test_gpu_knn_adds.py tests/test_gpu_pose_math.py tests/test_gpu_icp.py; do
  timeout 600 python -m pytest $t -x -q -m gpu 2>&1 | tail -25 >> gpurun_out/first.log
done
timeout 300 python tools/gpu_diag.py simt >> gpurun_out/first.log 2>&1
timeout 300 python tools/gpu_diag.py tc >> gpurun_out/first.log 2>&1
echo "diag tc exit: $?" >> gpurun_out/first.log
timeout 900 python -m pytest tests/test_gpu_net.py -x -q -m gpu 2>&1 | tail -40 >> gpurun_out/first.log
tail -150 gpurun_out/first.log
