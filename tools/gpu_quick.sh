#!/bin/bash
# quick GPU pass: parity tests + one bench line (value / frac / layers)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --steps 100 --warmup 5 $BENCH_ARGS > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err || tail -5 gpurun_out/bench_quick.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
r = d['roofline']
print('value %.0f e2e %.0f frac %.4f gemm_ms %.4f all_ms %.4f' % (d['value'], d['e2e']['value'], r['frac'], r['gemm_ms_per_step'], r['all_kernels_ms_per_step']))
print({k: round(v['ms_per_step'], 4) for k, v in r['layers'].items()})
print({k: round(v, 4) for k, v in r['other_kernels_ms_per_step'].items()})
t = (d.get('extra') or {}).get('refiner_training')
if t: print('train', t.get('objects_per_s'), t.get('ms_per_step'), t.get('error'))
PY
