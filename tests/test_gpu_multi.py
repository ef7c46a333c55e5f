"""GPU, 2 ranks (skipped on a one-GPU box; run with `gpurun --gpus 2`): the data-parallel refiner step on the kernels --
all-reduced flat gradient of two half-batches == whole-batch gradient (the NCCL sum is the only collective of the scope)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from autoposeestimation_b200 import synthetic as synth
from autoposeestimation_b200.densefusion.train_refiner import RefinerTrainer
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
nobj, B, N, M = 3, 8, 256, 200
rng = np.random.RandomState(5)
pts = torch.from_numpy((rng.randn(B, N, 3) * 0.05).astype(np.float32)).cuda()
emb = torch.from_numpy(rng.randn(B, 32, N).astype(np.float32)).cuda()
idx = torch.from_numpy(rng.randint(0, nobj, (B,)).astype(np.int64)).cuda()
model = torch.from_numpy(((rng.rand(B, M, 3) - 0.5) * 0.2).astype(np.float32)).cuda()
target = model + 0.01
sd = synth.refiner_state_dict(77, nobj)
sd['conv3_r.bias'] = sd['conv3_r.bias'].copy(); sd['conv3_r.bias'][0::4] += 1.0
whole = RefinerTrainer(sd, nobj, B, N, sym_list=[1])
whole.zero_grad(); whole.accumulate(pts, emb, idx, target, model)
lo, hi = rank * B // world, (rank + 1) * B // world
part = RefinerTrainer(sd, nobj, B, N, sym_list=[1])
part.zero_grad(); part.accumulate(pts[lo:hi], emb[lo:hi], idx[lo:hi], target[lo:hi], model[lo:hi])
part.allreduce_gradient()
rel = float((part.h.grads - whole.h.grads).norm() / whole.h.grads.norm())
part.optimizer_step(); whole.optimizer_step()
drift = float((part.h.params - whole.h.params).abs().max())
# the overlapped all-reduce (tail block of the gradient on a side stream while the backward pass finishes, head block after
# it) leaves the same gradient as the single all-reduce, block by block, up to the run-to-run noise of the fp32 atomics in
# the split-K weight gradients (measured here between two serial steps)
ta = RefinerTrainer(sd, nobj, B, N, sym_list=[1]); tb = RefinerTrainer(sd, nobj, B, N, sym_list=[1]); tc = RefinerTrainer(sd, nobj, B, N, sym_list=[1])
half = (pts[lo:hi], emb[lo:hi], idx[lo:hi], target[lo:hi], model[lo:hi])
ta.train_step(*half, overlap=True); tb.train_step(*half, overlap=False); tc.train_step(*half, overlap=False)
torch.cuda.synchronize()
bulk = ta.h.wait_bulk(torch.cuda.Stream())
def rel_blocks(x, y):
    return [float((x[a:b] - y[a:b]).norm() / y[a:b].norm()) for a, b in ((0, bulk), (bulk, x.numel()))]
r_ov, r_noise = rel_blocks(ta.h.grads, tb.h.grads), rel_blocks(tc.h.grads, tb.h.grads)
for it in range(2):                                        # and the parameters stay together over further steps
    ta.train_step(*half, overlap=True); tb.train_step(*half, overlap=False)
torch.cuda.synchronize()
pdrift = float((ta.h.params - tb.h.params).abs().max())
print('RESULT rank %d rel %.3e drift %.3e overlap %s noise %s pdrift %.2e bulk %d of %d'
      % (rank, rel, drift, r_ov, r_noise, pdrift, bulk, ta.h.grads.numel()), flush=True)
dist.barrier(); dist.destroy_process_group()
assert rel < 2e-3 and drift < 1e-5, (rel, drift)
assert all(a < max(1e-5, 10 * n) for a, n in zip(r_ov, r_noise)), (r_ov, r_noise)
assert pdrift < 6.5e-4                                     # 3 Adam steps of lr 1e-4: |step| <= lr, opposite signs at worst
assert 0 < bulk < ta.h.grads.numel() // 4
'''


def test_two_rank_allreduced_gradient_equals_whole_batch(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (gpurun --gpus 2)')
    script = tmp_path / 'worker.py'
    script.write_text(_WORKER)
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                        '--master-port', '29617', str(script), ROOT], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count('RESULT') == 2


def test_kernels_follow_the_tensors_device():
    """Handles and launches follow the device of their tensors, not the caller's current device (per-device kernel
    attributes, per-device SM count): the pipeline on cuda:1 while cuda:0 is current gives the bits of cuda:0."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (gpurun --gpus 2)')
    import numpy as np
    from autoposeestimation_b200 import ops, synthetic as synth
    nobj, B, N = 3, 20, 300
    sd_e, sd_r = synth.posenet_state_dict(5, nobj), synth.refiner_state_dict(1005, nobj)
    inp = synth.posenet_inputs(5, N, (40, 60), nobj, batch=B)
    outs = []
    for dev in ('cuda:0', 'cuda:1'):
        torch.cuda.set_device(0)                                  # the current device stays cuda:0 throughout
        est = ops.NetHandle(ops.NET_POSENET, sd_e, nobj, B, N, device=dev)
        ref = ops.NetHandle(ops.NET_REFINER, sd_r, nobj, B, N, device=dev)
        d = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in inp]
        poses, wm = ops.pose_pipeline(est, ref, *d)
        assert poses.device == torch.device(dev)
        src = torch.rand((20000, 3), dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(1)) * 80
        vox, cnt = ops.voxel_down_sample(src, torch.tensor([0, 20000], dtype=torch.int32, device=dev), 2.0)
        outs.append((poses.cpu(), int(cnt[0])))
    assert torch.equal(outs[0][0], outs[1][0]) and outs[0][1] > 0 and outs[1][1] > 0
