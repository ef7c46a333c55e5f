"""Development A/B of the ICP kernel builds on 3552 distinct registrations: python tools/icp_ab.py"""
import os, sys, subprocess, json
if len(sys.argv) == 1:
    for env in ({'APE_ICP_SMEM': '0'}, {'APE_ICP_SMEM': '-8'}):
        r = subprocess.run([sys.executable, __file__, 'run'], env=dict(os.environ, **env), capture_output=True, text=True)
        print(env, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:], flush=True)
    sys.exit(0)
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autoposeestimation_b200 import ops, synthetic as synth
dev = torch.device('cuda')
nreg, L = 3552, 5
scene = synth.Scene(3); nfr = (nreg + L - 1) // L
poses = scene.camera_poses(500, nfr)
lab = torch.empty((nfr, 480, 640), dtype=torch.uint8, device=dev); dep = torch.empty((nfr, 480, 640), dtype=torch.int16, device=dev)
for c0 in range(0, nfr, 128):
    lab[c0:c0 + 128], dep[c0:c0 + 128] = scene.render(poses[c0:c0 + 128], seed=11 + c0, device=dev)
cam = torch.tensor([[320., 240., 615., 615.]], dtype=torch.float64, device=dev).repeat(nfr, 1)
o = ops.surface_backproject_multi(lab, dep, cam, torch.from_numpy(poses).to(dev), [1, 2, 3, 4, 5], total_capacity=nfr * L * 8192)
vox, vc = ops.voxel_down_sample(o['points'], o['offsets'], 2.0, max_cloud_points=10240)
so = o['offsets'][:nreg + 1].contiguous(); vc = vc[:nreg].contiguous()
tgt = torch.from_numpy(np.concatenate([scene.models_pert[v % L] for v in range(nreg)])).to(dev)
to = torch.arange(0, nreg + 1, device=dev, dtype=torch.int32) * 2000
T, info = ops.icp_p2p(vox, so, tgt, to, 10.0, src_count=vc)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    T, info = ops.icp_p2p(vox, so, tgt, to, 10.0, src_count=vc)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print('%.0f reg/s  %.3f ms  checksum %.12f iters %.3f' % (nreg / ms * 1e3, ms, float(T.sum()), float(info[:, 2].mean())))
