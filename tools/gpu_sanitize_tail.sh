#!/bin/bash
# compute-sanitizer over the one-launch refiner tail (gemm_tail.cuh) and the overlapped training step. Output: gpurun_out/sanitize_r02b.log
mkdir -p gpurun_out
L=gpurun_out/sanitize_r02b.log
: > $L
run() {
  echo "== $1 :: $2" >> $L
  eval timeout 900 compute-sanitizer --tool $1 --report-api-errors no --error-exitcode 77 --print-limit 5 python -m pytest $2 -x -q 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard|error|Error" | head -12 >> $L
}
run memcheck "tests/test_gpu_net.py -k 'tcgen05_matches_simt and (130 or 17)'"
run memcheck "tests/test_gpu_net.py -k 'pose_pipeline_vs_oracle'"
run racecheck "tests/test_gpu_net.py -k 'tcgen05_matches_simt and 17'"
run memcheck "tests/test_gpu_train.py -k 'forward_backward_vs_fp32 and 3-100'"
cat $L
