#!/bin/bash
# GPU pass for the training path: tests, bench (with the training leg), launch list of one training step
mkdir -p gpurun_out
L=gpurun_out/train_round.log
echo "== pytest train + dropin" > $L
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_dropin.py -x -q 2>&1 | tail -15 >> $L
echo "== bench" >> $L
timeout 900 python bench.py --steps 50 --warmup 5 --no-icp > gpurun_out/bench_train.json 2>> $L
python - >> $L <<'PY'
import json
d = json.load(open('gpurun_out/bench_train.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
print(json.dumps(d['extra'], indent=1))
PY
tail -80 $L
