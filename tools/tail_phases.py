"""Developer aid: clock64 stamps of the phases of refiner_tail_kernel (gemm_tail.cuh), median / max over the CTAs of one launch.
Needs a TIMING build of the library: add `#define APE_TAIL_TIMING` in front of `#include "gemm_tail.cuh"` in csrc/net.cu,
rebuild (python -m autoposeestimation_b200.build), run this on the GPU box, then remove the define again (the timing build
exports an extra symbol, ape_debug_tail, that include/ape_b200.h does not declare).
Stamps: 0 start, 1 weights issued, 2 after the dependency wait, 3 pooled operand built, 4 MMAs retired, 5 partial tile written,
6 barrier, 7 conv1 reduced, 8 conv2 partials, 9 barrier, 10 conv2 reduced, 11 conv3 partials pushed, 12 barrier, 13 done."""
import ctypes, sys, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
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
