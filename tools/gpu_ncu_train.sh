#!/bin/bash
# ncu: launch list of training steps + full capture of one refinement iteration of the second step (tensor-core and
# streaming kernels).  The .ncu-rep is converted to CSV on the box; only small files travel back.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/train_launches.csv \
    python tools/train_once.py 256 2 > gpurun_out/ncu_train_list.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'dy6|conv1_wgrad|gemm_bf16_bwd|gemm_split|refine_loss|sgemm_small|fold_partials|dense_batch|frontend' \
    -s 38 -c 19 -o /tmp/prof_train -f python tools/train_once.py 256 2 > gpurun_out/ncu_train.log 2>&1
ncu -i /tmp/prof_train.ncu-rep --page raw --csv > gpurun_out/prof_train_raw.csv 2>/dev/null
tail -2 gpurun_out/ncu_train.log
ls -la gpurun_out/
