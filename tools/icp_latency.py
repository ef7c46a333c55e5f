"""Latency of ONE registration against the iteration cap (setup cost vs per-iteration cost): python tools/icp_latency.py"""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from autoposeestimation_b200 import ops, synthetic as synth
rng = np.random.RandomState(1)
tgt = synth.ellipsoid_cloud(rng, 2000)
R = synth.random_rotation(rng, 0.15); t = np.array([4.0, -3.0, 2.0])
src = (tgt[rng.choice(2000, 900, replace=False)] @ R.T + t)
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
S, Tg = d(src), d(tgt)
so, to = d(np.array([0, len(src)], np.int32)), d(np.array([0, len(tgt)], np.int32))
for mi in (0, 1, 2, 5, 10, 20, 40):
    for _ in range(3):
        T, info = ops.icp_p2p(S, so, Tg, to, 10.0, rel_fitness=0.0, rel_rmse=0.0, max_iter=mi)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        T, info = ops.icp_p2p(S, so, Tg, to, 10.0, rel_fitness=0.0, rel_rmse=0.0, max_iter=mi)
    e1.record(); torch.cuda.synchronize()
    print('max_iter %2d: %7.1f us per registration, iterations run %d, corr %d' % (mi, e0.elapsed_time(e1) * 100, int(info[0, 2]), int(info[0, 3])))
