#!/bin/bash
# compute-sanitizer over the parity tests (small shapes): memcheck on every kernel family incl. the tcgen05 / TMA / cluster
# kernels, racecheck on the shared-memory heavy SIMT kernels.  Output: gpurun_out/sanitize_r02.log
# (--report-api-errors no: torch itself probes cuKernelGetFunction with handles that fail by design; those are not ours)
mkdir -p gpurun_out
L=gpurun_out/sanitize_r02.log
: > $L
run() {  # tool, pytest selection
  echo "== $1 :: $2" >> $L
  eval timeout 900 compute-sanitizer --tool $1 --report-api-errors no --error-exitcode 77 --print-limit 5 python -m pytest $2 -x -q 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard|error|Error" | head -12 >> $L
}
run memcheck "tests/test_gpu_net.py -k 'golden and case0 and (tcgen05-0 or tcgen05]) or pass_table'"
run memcheck "tests/test_gpu_gather.py -k 'gather_emb or layouts'"
run memcheck "tests/test_gpu_icp.py -k 'c1_vs_oracle or kabsch or large_clouds and 40000'"
run memcheck "tests/test_gpu_backproject.py -k 'multi_label or mask_bbox_choose'"
run memcheck "tests/test_gpu_dropin.py -k 'loss_forward_and_gradient or nonsymmetric'"
run memcheck "tests/test_gpu_train.py -k 'refine_loss or accumulates'"
run memcheck "tests/test_gpu_filters.py -k 'ragged or radius_outlier_vs_oracle or mahalanobis'"
run racecheck "tests/test_gpu_icp.py -k 'c1_vs_oracle and 0 or large_clouds and 16385'"
run racecheck "tests/test_gpu_dropin.py -k 'loss_forward_and_gradient'"
run racecheck "tests/test_gpu_backproject.py -k 'multi_label'"
run racecheck "tests/test_gpu_knn_adds.py -k 'add_metric_vs_oracle'"
cat $L
