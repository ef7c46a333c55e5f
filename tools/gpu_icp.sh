#!/bin/bash
# ICP: parity tests + the bench's label-path leg
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_icp.py tests/test_gpu_filters.py tests/test_gpu_dropin.py -x -q -m gpu 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 40 --warmup 5 --no-train > gpurun_out/bench_icp.json 2> gpurun_out/bench_icp.err || tail -5 gpurun_out/bench_icp.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_icp.json').read().strip().splitlines()[-1])
e = d['extra']
print('bp', {k: v for k, v in e['backprojection'].items() if k not in ('kernels', 'roofline')}, e['backprojection']['roofline']['frac'])
print('icp', {k: v for k, v in e['icp'].items() if k != 'roofline'})
PY
