"""Development diagnostics: where does the host-buffer Runner spend its time?  python tools/e2e_diag.py [steps]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autoposeestimation_b200 import ops, synthetic as synth  # noqa: E402
from autoposeestimation_b200.densefusion.estimate_poses import Runner  # noqa: E402

B, N, CROP, NOBJ = 64, 500, (120, 160), 5
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
est = ops.NetHandle(ops.NET_POSENET, synth.posenet_state_dict(7, NOBJ), NOBJ, B, N)
ref = ops.NetHandle(ops.NET_REFINER, synth.refiner_state_dict(1007, NOBJ), NOBJ, B, N)
sets = [[torch.from_numpy(a).pin_memory() for a in synth.posenet_inputs(10 + i, N, CROP, NOBJ, batch=B)] for i in range(3)]
dsets = [[t.cuda() for t in s] for s in sets]
poses = torch.empty((B, 7), dtype=torch.float64, device='cuda')
for i in range(5):
    ops.pose_pipeline(est, ref, *dsets[i % 3], out=poses)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    ops.pose_pipeline(est, ref, *dsets[i % 3], out=poses)
e1.record(); torch.cuda.synchronize()
print('device-resident: %.4f ms/step' % (e0.elapsed_time(e1) / steps))
for name, kw in (('auto', {}), ('auto 8thr', dict(host_threads=8)), ('auto 4thr', dict(host_threads=4)), ('auto 2thr', dict(host_threads=2)), ('zc 1.0', dict(zero_copy_fraction=1.0))):
    r = Runner(est, ref, B, N, CROP[0] * CROP[1], **kw)
    for i in range(6):
        r.submit(*sets[i % 3])
    r.drain()
    r.cpu_ms = {}
    t0 = time.perf_counter()
    e0.record()
    for i in range(steps):
        r.submit(*sets[i % 3])
    t_sub = time.perf_counter() - t0
    r.drain()
    e1.record(); torch.cuda.synchronize()
    print('%-14s %.4f ms/step (cpu submit %.4f ms/step) cal=%s cpu phases: %s' % (
        name, e0.elapsed_time(e1) / steps, t_sub / steps * 1e3, r.calibration and {k: (round(v, 3) if not isinstance(v, dict) else {a: round(b, 3) for a, b in v.items()}) for k, v in r.calibration.items()},
        {k: round(v / steps, 4) for k, v in r.cpu_ms.items()}), flush=True)
