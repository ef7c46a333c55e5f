#!/bin/bash
# compute-sanitizer memcheck over the parity tests of the newer kernels (small shapes)
mkdir -p gpurun_out
L=gpurun_out/sanitize.log
: > $L
for t in "tests/test_gpu_train.py -k 'refine_loss or accumulates or 3-100'" "tests/test_gpu_filters.py -k 'ragged or radius_outlier_vs_oracle or mahalanobis'" \
         "tests/test_gpu_backproject.py -k 'mask_bbox_choose'" "tests/test_gpu_dropin.py -k 'fused_forward or nonsymmetric'"; do
  echo "== $t" >> $L
  eval timeout 900 compute-sanitizer --tool memcheck --error-exitcode 77 --print-limit 5 python -m pytest $t -x -q 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error|Error" | head -12 >> $L
done
cat $L
