#!/bin/bash
# label path (option 4): parity tests, bench extra leg, ncu of back-projection / voxel / ICP kernels
mkdir -p gpurun_out
L=gpurun_out/label.log
nvidia-smi -L > $L 2>&1
timeout 900 python -m pytest tests/test_gpu_backproject.py tests/test_gpu_icp.py tests/test_gpu_dropin.py tests/test_gpu_net.py -x -q -m gpu 2>&1 | tail -15 >> $L
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_label.json 2>> $L
python - >> $L <<'PY'
import json
d = json.loads(open('gpurun_out/bench_label.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms/step', d['ms_per_step'])
print(json.dumps(d['extra'], indent=1))
PY
python tools/bench_layers.py gpurun_out/bench_label.json >> $L 2>&1
if [ "$1" != "noprof" ]; then
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'icp|voxel|surface' -c 8 -o gpurun_out/prof_label -f \
    python bench.py --steps 4 --warmup 1 > gpurun_out/ncu_label.log 2>&1
tail -3 gpurun_out/ncu_label.log >> $L
fi
tail -${TAIL:-100} $L
