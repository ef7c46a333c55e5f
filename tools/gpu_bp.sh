#!/bin/bash
# back-projection / label path: parity tests + the bench's label-path leg
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backproject.py tests/test_gpu_filters.py tests/test_gpu_dropin.py -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --steps 100 --warmup 5 --no-train > gpurun_out/bench_bp.json 2> gpurun_out/bench_bp.err || tail -5 gpurun_out/bench_bp.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_bp.json').read().strip().splitlines()[-1])
e = d['extra']
print('value %.0f' % d['value'])
print('bp', {k: v for k, v in e['backprojection'].items() if k != 'kernels'})
print('icp', e['icp']['registrations_per_s'], e['icp']['ms_per_launch'])
PY
