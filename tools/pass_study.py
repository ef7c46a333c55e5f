"""Per-layer error budget of the split-bf16 GEMM (development + evidence for DESIGN.md 5): max |delta pose| and arg-max
flips against the fp32 reference restatement (oracle, torch-CPU) over seeded objects, for 3 / 2 / 1 bf16 products per
layer.   python tools/pass_study.py [n_objects]  ->  gpurun_out/pass_study.json"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autoposeestimation_b200 import ops, synthetic as synth  # noqa: E402
from oracle import densefusion as odf, pose_math as pm  # noqa: E402

NOBJ, N, CROP = 5, 500, (120, 160)
n_obj = int(sys.argv[1]) if len(sys.argv) > 1 else 256
B = 64
torch.set_num_threads(os.cpu_count())
results = {}
all_ref = []
nets = []
for wseed in (7, 8):                                          # two weight sets
    sd_e, sd_r = synth.posenet_state_dict(wseed, NOBJ), synth.refiner_state_dict(1000 + wseed, NOBJ)
    te, tr = synth.to_torch(sd_e), synth.to_torch(sd_r)
    est = ops.NetHandle(ops.NET_POSENET, sd_e, NOBJ, B, N); ref = ops.NetHandle(ops.NET_REFINER, sd_r, NOBJ, B, N)
    batches = []
    for k in range(n_obj // (2 * B)):
        inp = synth.posenet_inputs(100 * wseed + k, N, CROP, NOBJ, batch=B)
        want = []
        with torch.no_grad():
            for b in range(B):
                t = [torch.from_numpy(a[b:b + 1]) for a in inp]
                res = odf.canonical_prediction(te, tr, t[0], t[1], t[2], t[3], NOBJ, iterations=2)
                want.append(np.concatenate([res['q'], res['t']]))
        batches.append(([torch.from_numpy(a).cuda() for a in inp], np.stack(want)))
    nets.append((est, ref, batches))
print('oracle poses done', flush=True)


def run(pn, rf):
    dq, dt, flips = [], [], 0
    for est, ref, batches in nets:
        est.set_passes(pn); ref.set_passes(rf)
        for d, want in batches:
            poses, wm = ops.pose_pipeline(est, ref, *d, iterations=2, canonical=True)
            poses = poses.cpu().numpy()
            for b in range(B):
                dq.append(pm.rotation_angle_between(poses[b, :4], want[b, :4]))
                dt.append(np.abs(poses[b, 4:] - want[b, 4:]).max())
    dq, dt = np.array(dq), np.array(dt)
    return dict(max_rot_rad=float(dq.max()), max_trans_m=float(dt.max()), n_rot_over_1e3=int((dq > 1e-3).sum()),
                n_trans_over_1e4=int((dt > 1e-4).sum()), p99_rot=float(np.percentile(dq, 99)), p99_trans=float(np.percentile(dt, 99)))


F = [7] * 6
cfgs = {'all 3-pass': (F, F)}
for li, lname in enumerate(['conv2', 'conv5', 'conv6', 'heads1', 'heads2', 'heads3']):
    for m, mname in ((6, 'drop A_lo'), (5, 'drop W_lo'), (4, 'bf16')):
        pn = list(F); pn[li] = m
        cfgs['PN %s %s' % (lname, mname)] = (pn, F)
for li, lname in enumerate(['conv2', 'conv5', 'conv6']):
    for m, mname in ((6, 'drop A_lo'), (5, 'drop W_lo'), (4, 'bf16')):
        rf = list(F); rf[li] = m
        cfgs['RF %s %s' % (lname, mname)] = (F, rf)
cfgs['RF all drop A_lo'] = (F, [6, 6, 6, 7, 7, 7])
cfgs['RF all drop A_lo + PN conv6 drop A_lo'] = ([7, 7, 6, 7, 7, 7], [6, 6, 6, 7, 7, 7])
cfgs['RF all drop A_lo + PN conv5,6 drop A_lo'] = ([7, 6, 6, 7, 7, 7], [6, 6, 6, 7, 7, 7])
cfgs['RF all bf16'] = (F, [4, 4, 4, 7, 7, 7])
cfgs['RF conv5,6 bf16 conv2 drop A_lo'] = (F, [6, 4, 4, 7, 7, 7])
cfgs['RF c2,c5 dropA c6 bf16 + PN conv6 drop A_lo'] = ([7, 7, 6, 7, 7, 7], [6, 6, 4, 7, 7, 7])
cfgs['everything drop A_lo'] = ([6] * 6, [6] * 6)
for name, (pn, rf) in cfgs.items():
    results[name] = dict(run(pn, rf), pn=pn, rf=rf)
    print('%-44s %s' % (name, {k: (('%.2e' % v) if isinstance(v, float) else v) for k, v in results[name].items() if k not in ('pn', 'rf')}), flush=True)
os.makedirs('gpurun_out', exist_ok=True)
json.dump(dict(n_objects=len(nets) * len(nets[0][2]) * B, gates=dict(rot_rad=1e-3, trans_m=1e-4), results=results),
          open('gpurun_out/pass_study.json', 'w'), indent=1)
