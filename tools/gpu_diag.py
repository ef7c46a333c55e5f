"""Development diagnostics on the GPU box: error statistics per stage and quick timings.
Usage: python tools/gpu_diag.py [simt|tc|all]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autoposeestimation_b200 import ops  # noqa: E402
from oracle import densefusion as odf, synth  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main(which):
    print('device', torch.cuda.get_device_name(0), flush=True)
    nobj, B, N = 5, 64, 500
    sd_e = synth.posenet_state_dict(1, nobj); sd_r = synth.refiner_state_dict(1001, nobj)
    est = ops.NetHandle(ops.NET_POSENET, sd_e, nobj, B, N)
    ref = ops.NetHandle(ops.NET_REFINER, sd_r, nobj, B, N)
    out_img, cloud, choose, idx = synth.posenet_inputs(1, N, (120, 160), nobj, batch=B)
    d = [dev(a) for a in (out_img, cloud, choose, idx)]
    te, tr = synth.to_torch(sd_e), synth.to_torch(sd_r)
    with torch.no_grad():
        o = odf.posenet_geometry(te, torch.from_numpy(out_img[:1]), torch.from_numpy(cloud[:1]), torch.from_numpy(choose[:1]),
                                 torch.from_numpy(idx[:1]), nobj)
    impls = [('simt', ops.GEMM_SIMT), ('tc', ops.GEMM_TCGEN05)]
    if which != 'all':
        impls = [i for i in impls if i[0] == which]
    res = {}
    for name, gi in impls:
        est.set_gemm(gi); ref.set_gemm(gi)
        r, t, c, emb = est.posenet_forward(*d)
        torch.cuda.synchronize()
        print('[%s] posenet vs oracle (object 0): max|dr|=%.3e (scale %.2f)  max|dt|=%.3e (scale %.2f)  max|dc|=%.3e' % (
            name, float((r[0].cpu() - o[0][0]).abs().max()), float(o[0].abs().max()),
            float((t[0].cpu() - o[1][0]).abs().max()), float(o[1].abs().max()), float((c[0].cpu() - o[2][0]).abs().max())), flush=True)
        res[name] = (r, t, c)
        ms = timeit(lambda: est.posenet_forward(*d))
        print('[%s] posenet_forward B=64 N=500: %.3f ms' % (name, ms), flush=True)
        ms = timeit(lambda: ops.pose_pipeline(est, ref, *d, iterations=2, canonical=True))
        print('[%s] pose_pipeline (PoseNet + 2 refine) B=64: %.3f ms -> %.0f frames/s' % (name, ms, B / ms * 1e3), flush=True)
    if len(res) == 2:
        for k, nm in enumerate(('r', 't', 'c')):
            print('tc vs simt max|d%s| = %.3e' % (nm, float((res['simt'][k] - res['tc'][k]).abs().max())))


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'all')
