import ctypes, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from autoposeestimation_b200 import ops, _lib, synthetic as synth


lib = _lib.load()
B, N = 64, 500
sd_e = synth.posenet_state_dict(1, 5); sd_r = synth.refiner_state_dict(2, 5)
est = ops.NetHandle(ops.NET_POSENET, sd_e, 5, B, N); ref = ops.NetHandle(ops.NET_REFINER, sd_r, 5, B, N)
img, cloud, choose, idx = synth.posenet_inputs(5, N, (120, 160), 5, batch=B)
d = [torch.from_numpy(a).cuda() for a in (img, cloud, choose, idx)]
for _ in range(5):
    out = ops.pose_pipeline(est, ref, d[0], d[1], d[2], d[3], iterations=2, canonical=True)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (64 * 16))()
print('rc', lib.ape_debug_tail(buf))
a = np.array(buf[:], dtype=np.int64).reshape(64, 16)[:32, :14]
d = a - a[:, :1]
np.set_printoptions(linewidth=250)
print('phase stamps (cycles from CTA start), median over CTAs:'); print(np.median(d, axis=0).astype(int))
print('max:'); print(d.max(axis=0))
print('start skew (cycles):', (a[:, 0] - a[:, 0].min())[:32])
