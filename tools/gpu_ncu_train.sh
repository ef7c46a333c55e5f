#!/bin/bash
# ncu: launch list of one training step + full capture of the tensor-core / streaming training kernels (one iteration
# of the second step).  The .ncu-rep is converted to CSV on the box; only small files travel back.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/train_launches.csv \
    python tools/train_once.py 256 2 > gpurun_out/ncu_train_list.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'dy6|conv1_wgrad|gemm_bf16_bwd|gemm_split' -s 22 -c 11 -o /tmp/prof_train -f \
    python tools/train_once.py 256 2 > gpurun_out/ncu_train.log 2>&1
ncu -i /tmp/prof_train.ncu-rep --page raw --csv > gpurun_out/prof_train_raw.csv 2>/dev/null
ls -la /tmp/prof_train.ncu-rep gpurun_out/ | tail -8
sz=$(stat -c %s /tmp/prof_train.ncu-rep); if [ "$sz" -lt 30000000 ]; then cp /tmp/prof_train.ncu-rep gpurun_out/; fi
tail -3 gpurun_out/ncu_train.log
