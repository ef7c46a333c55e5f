"""Diagnostic: per-tensor gradient error of the bf16 training backward vs fp32 autograd (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import densefusion as odf
sys.path.insert(0, 'tests')
import test_gpu_train as T
from autoposeestimation_b200 import ops

B, N, nobj = int(sys.argv[1]) if len(sys.argv) > 1 else 5, int(sys.argv[2]) if len(sys.argv) > 2 else 200, 3
sd_np = T._state_dict(21, nobj)
points, emb, idx, _, _ = T._inputs(5 + B, B, N, 16, nobj)
rng = np.random.RandomState(9)
d_r = rng.randn(B, 4).astype(np.float32); d_t = rng.randn(B, 3).astype(np.float32)
tr = ops.RefinerTrainerHandle(sd_np, nobj, B, N)
r2, t2 = tr.forward(T._dev(points), T._dev(emb), T._dev(idx))
tr.backward(T._dev(points), T._dev(emb), T._dev(idx), T._dev(d_r), T._dev(d_t))
torch.cuda.synchronize()
for mode in ('fp32', 'bf16w'):
    sd = T._ref_sd(sd_np)
    if mode == 'bf16w':      # reference with bf16-rounded trunk weights (what the tensor cores see)
        for k in ('feat.conv2.weight', 'feat.e_conv2.weight', 'feat.conv5.weight', 'feat.conv6.weight'):
            sd[k] = sd[k].detach().bfloat16().float().requires_grad_(True)
    for b in range(B):
        r, t = odf.refiner_forward(sd, torch.from_numpy(points[b:b + 1]), torch.from_numpy(emb[b:b + 1]),
                                   torch.from_numpy(idx[b:b + 1]).view(1, 1), nobj)
        ((r[0] * torch.from_numpy(d_r[b])).sum() + (t[0] * torch.from_numpy(d_t[b])).sum()).backward()
    print('== reference:', mode)
    for key in tr.table:
        g = tr.view(key, tr.grads).cpu()
        rel, cos = T._rel_cos(g, sd[key].grad.reshape(g.shape))
        print('%-22s rel %.4f cos %.6f  |ref| %.3e' % (key, rel, cos, float(sd[key].grad.norm())))
