"""Print the per-layer table of a bench.py JSON line: python tools/bench_layers.py gpurun_out/bench.json"""
import json
import sys

for path in sys.argv[1:]:
    txt = open(path).read().strip().splitlines()
    if not txt:
        print(path, 'EMPTY'); continue
    d = json.loads(txt[-1])
    r = d['roofline']
    print('%s: value %.0f %s  ms/step %.4f  e2e %.0f  frac %.3f (executed %.3f)  gemm %.4f ms  all kernels %.4f ms' % (
        path, d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], r['frac'], r.get('executed_frac', 0), r['gemm_ms_per_step'],
        r['all_kernels_ms_per_step']))
    for k, v in sorted(r['layers'].items()):
        print('   %-16s %5.1f launches  %8.4f ms  alg %7.1f TF  exec %7.1f TF' % (k, v['launches_per_step'], v['ms_per_step'],
                                                                              v['algorithmic_tflops'], v['executed_bf16_tflops']))
    for k, v in sorted(r['other_kernels_ms_per_step'].items()):
        print('   %-16s %8.4f ms' % (k, v))
    print('   clocks', d.get('clocks'))
