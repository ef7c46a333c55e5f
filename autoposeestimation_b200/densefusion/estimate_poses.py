"""Batched option-6 geometry block from HOST buffers (the call a user of the drop-in makes).

`Runner` double-buffers: the H2D copy of batch i+1 (pinned host memory, copy stream) overlaps the
kernels of batch i; the poses of every batch are copied back to pinned host memory.  No CPU fallback.
Replaces the per-object loop of pipeline/utils.py:556-571 (three H2D + three D2H syncs per object).
"""
import numpy as np
import torch

from .. import ops


class Runner:
    def __init__(self, estimator, refiner, max_batch, n_points, crop_pixels, iterations=2, canonical=True, device=None):
        self.est, self.ref = estimator, refiner
        self.iterations, self.canonical = iterations, canonical
        dev = device or torch.device('cuda', torch.cuda.current_device())
        self.dev = dev
        self.copy_stream = torch.cuda.Stream(device=dev)
        mk = lambda shape, dt: [torch.empty(shape, dtype=dt, device=dev) for _ in range(2)]
        self.d_img = mk((max_batch, 32, crop_pixels), torch.float32)
        self.d_cloud = mk((max_batch, n_points, 3), torch.float32)
        self.d_choose = mk((max_batch, n_points), torch.int64)
        self.d_idx = mk((max_batch,), torch.int64)
        self.d_pose = mk((max_batch, 7), torch.float64)
        self.h_pose = [torch.empty((max_batch, 7), dtype=torch.float64).pin_memory() for _ in range(2)]
        self.copied = [torch.cuda.Event() for _ in range(2)]
        self.done = [torch.cuda.Event() for _ in range(2)]
        self.i = 0
        self.last = None

    def submit(self, out_img, cloud, choose, idx):
        """Host tensors (pinned for true async): out_img [B,32,H,W] fp32, cloud [B,N,3] fp32,
        choose [B,1,N] int64, idx [B,1] int64.  Returns immediately."""
        s = self.i % 2
        B = cloud.shape[0]
        main = torch.cuda.current_stream(self.dev)
        if self.i >= 2:
            self.copy_stream.wait_event(self.done[s])         # device slot s is free again
        with torch.cuda.stream(self.copy_stream):
            self.d_img[s][:B].copy_(out_img.reshape(B, 32, -1), non_blocking=True)
            self.d_cloud[s][:B].copy_(cloud, non_blocking=True)
            self.d_choose[s][:B].copy_(choose.reshape(B, -1), non_blocking=True)
            self.d_idx[s][:B].copy_(idx.reshape(B), non_blocking=True)
            self.copied[s].record(self.copy_stream)
        main.wait_event(self.copied[s])
        ops.pose_pipeline(self.est, self.ref, self.d_img[s][:B], self.d_cloud[s][:B], self.d_choose[s][:B], self.d_idx[s][:B],
                          iterations=self.iterations, canonical=self.canonical, out=self.d_pose[s][:B])
        self.h_pose[s][:B].copy_(self.d_pose[s][:B], non_blocking=True)
        self.done[s].record(main)
        self.last = (s, B)
        self.i += 1

    def drain(self):
        """Wait for everything submitted; returns the poses [B,7] fp64 (wxyz, t) of the last batch (host)."""
        torch.cuda.synchronize(self.dev)
        if self.last is None:
            return None
        s, B = self.last
        return self.h_pose[s][:B].clone()


def estimate_poses(estimator, refiner, out_img, cloud, choose, idx, iterations=2, canonical=True):
    """One-shot convenience wrapper: numpy / CPU tensors in, numpy poses [B,7] (wxyz, t; metres) out."""
    t = [torch.as_tensor(np.ascontiguousarray(a)) if not isinstance(a, torch.Tensor) else a for a in (out_img, cloud, choose, idx)]
    B, N = t[1].shape[0], t[1].shape[1]
    r = Runner(estimator, refiner, B, N, int(np.prod(t[0].shape[2:])), iterations, canonical)
    r.submit(*t)
    return r.drain().numpy()
