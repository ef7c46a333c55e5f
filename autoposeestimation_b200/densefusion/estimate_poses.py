"""Batched option-6 geometry block from HOST buffers (the call a user of the drop-in makes).

Replaces the per-object loop of pipeline/utils.py:556-571 (three H2D + three D2H syncs per object).  `Runner`
double-buffers: the hand-over of batch i+1 overlaps the kernels of batch i; the poses of every batch are copied back to
pinned host memory.  No CPU fallback: the networks and the pose math only ever run on the sm_100a kernels.

Hand-over of the colour-encoder output.  The kernels need 32 channels at the N sampled pixels of every object
(network.py:100-102), i.e. 4.1 MB of a 157 MB [64,32,120,160] fp32 map.  Shipping the whole map (round 1) made the call
PCIe-bound (2.9 ms per batch against 0.76 ms of kernels).  `transfer=`
  'gather' (default)  the whole map never crosses the bus for all objects.  For a pinned [B,32,H,W] map the OBJECTS of a
                      batch are split three ways, by what each path costs (calibrated on the first batch):
                      (a) the host thread pool gathers the sampled columns into a pinned staging buffer, then one small
                          H2D (ops.host_gather_*: 0.54 ms per batch on 16 cores) -- costs HOST time;
                      (b) the copy engine ships some objects' whole maps (cudaMemcpyAsync, 2.8 ms per batch: neither
                          SMs nor cores) and a device gather picks the columns -- costs PCIe time only;
                      (c) the zero-copy gather kernel reads the map in place over PCIe (ops.gather_emb, 1.26 ms per
                          batch stand-alone) -- costs DEVICE time: about half of it shows up in the step.
                      All three read the same host DRAM and slow each other down (measured), so the calibration keeps
                      everything on the host pool while the host stage is shorter than the device time; otherwise it
                      MEASURES a few steps of the real pipeline with no / half / the modelled zero-copy share and keeps
                      the fastest; the copy-engine share is there for explicit use (`dma_fraction=`).
                      A channels_last map
                      (`t.contiguous(memory_format=torch.channels_last)`: one point = one 128-byte line) goes through the
                      zero-copy kernel alone (0.09 ms); a pageable (unpinned) map through the host pool alone.
  'full'              cudaMemcpyAsync of the whole map, gather in the front-end kernel (the round-1 path; also what a
                      DEVICE-resident map gets, without the copy).
The step's kernels are replayed from a CUDA graph per buffer slot, so the host thread spends its time in the gather, not
in ~40 launches.
"""
import time

import numpy as np
import torch

from .. import ops


class Runner:
    def __init__(self, estimator, refiner, max_batch, n_points, crop_pixels, iterations=2, canonical=True, device=None,
                 transfer='gather', zero_copy_fraction=None, host_threads=0, use_graph=True, dma_fraction=None):
        assert transfer in ('gather', 'full')
        self.est, self.ref = estimator, refiner
        self.iterations, self.canonical = iterations, canonical
        dev = device or torch.device('cuda', torch.cuda.current_device())
        self.dev = dev
        self.transfer, self.zc_fraction, self.host_threads, self.use_graph = transfer, zero_copy_fraction, host_threads, use_graph
        self.dma_fraction = dma_fraction if dma_fraction is not None else (0.0 if zero_copy_fraction is not None else None)
        self.crop_pixels = crop_pixels
        import os
        self.copy_stream = torch.cuda.Stream(device=dev, priority=int(os.environ.get('APE_RUNNER_PRIO', '-1')))
        mk = lambda shape, dt: [torch.empty(shape, dtype=dt, device=dev) for _ in range(2)]
        self.d_img = mk((max_batch, 32, crop_pixels), torch.float32) if transfer == 'full' else [None, None]   # 'gather': sized on demand
        self.d_emb = mk((max_batch, 32, n_points), torch.float32)
        self.d_cloud = mk((max_batch, n_points, 3), torch.float32)
        self.d_choose = mk((max_batch, n_points), torch.int64)
        self.d_idx = mk((max_batch,), torch.int64)
        self.d_pose = mk((max_batch, 7), torch.float64)
        self.h_pose = [torch.empty((max_batch, 7), dtype=torch.float64).pin_memory() for _ in range(2)]
        self.h_stage = [torch.empty((max_batch, 32, n_points), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.copied = [torch.cuda.Event() for _ in range(2)]
        self.done = [torch.cuda.Event() for _ in range(2)]
        self.graphs = {}
        self.i = 0
        self.last = None
        self.calibration = None
        self.cpu_ms = None                                    # set to {} to accumulate host-side time per phase (diagnostics)

    # ------------------------------------------------------------------------------------------------
    def calibrate(self, out_img, cloud, choose, idx, reps=3, trial_steps=16):
        """Time the two gather paths and the step's kernels on this batch and choose how many objects go through the
        zero-copy kernel, the copy engine and the host pool (see the module docstring).
        -> dict(zero_copy_ms, host_ms, dma_ms, compute_ms, zero_copy_fraction, dma_fraction, host_fraction, predicted_ms_per_step)."""
        B = out_img.shape[0]
        torch.cuda.synchronize(self.dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        with torch.cuda.stream(self.copy_stream):
            self.d_cloud[0][:B].copy_(cloud, non_blocking=True)
            self.d_idx[0][:B].copy_(idx.reshape(B), non_blocking=True)
            self.d_choose[0][:B].copy_(choose.reshape(B, -1), non_blocking=True)
            ops.gather_emb(out_img, self.d_choose[0][:B], out=self.d_emb[0][:B])        # warm-up (page tables, kernel load)
            ev[0].record(self.copy_stream)
            for _ in range(reps):
                ops.gather_emb(out_img, self.d_choose[0][:B], out=self.d_emb[0][:B])
            ev[1].record(self.copy_stream)
            run = lambda: ops.pose_pipeline(self.est, self.ref, self.d_emb[0][:B], self.d_cloud[0][:B], None, self.d_idx[0][:B],
                                            iterations=self.iterations, canonical=self.canonical, out=self.d_pose[0][:B], gathered=True)
            run()
            ev[2].record(self.copy_stream)
            for _ in range(reps):
                run()
            ev[3].record(self.copy_stream)
        ev[3].synchronize()
        t_zc, t_gpu = ev[0].elapsed_time(ev[1]) / reps, ev[2].elapsed_time(ev[3]) / reps
        ch = choose.reshape(B, -1).contiguous()
        ops.host_gather_begin(out_img, ch, self.h_stage[0], 0, B, self.host_threads); ops.host_gather_wait()
        t0 = time.perf_counter()
        for _ in range(reps):
            ops.host_gather_begin(out_img, ch, self.h_stage[0], 0, B, self.host_threads); ops.host_gather_wait()
        t_host = (time.perf_counter() - t0) / reps * 1e3
        # copy engine: whole maps of a few objects, scaled to the batch
        nb = max(1, min(8, B))
        tmp = torch.empty((nb, 32, self.crop_pixels), dtype=torch.float32, device=self.dev)
        with torch.cuda.stream(self.copy_stream):
            tmp.copy_(out_img[:nb].reshape(nb, 32, -1), non_blocking=True)
            ev[0].record(self.copy_stream)
            for _ in range(reps):
                tmp.copy_(out_img[:nb].reshape(nb, 32, -1), non_blocking=True)
            ev[1].record(self.copy_stream)
        ev[1].synchronize()
        t_dma = ev[0].elapsed_time(ev[1]) / reps * B / nb
        del tmp
        # Measured (tools/e2e_diag.py, profiles/r02_e2e_diag.txt): the copy engine and the zero-copy kernel both read the host
        # DRAM the pool is hammering, so adding either SLOWS the pool down (16 threads: host only 0.69 ms per step, +23 % of the
        # objects by copy engine 0.85 ms; 12 threads: 0.76 against 0.93) -- they only pay once the host is far behind the
        # device.  Policy: everything through the pool while the host stage is shorter than the device time; otherwise the
        # zero-copy share that lets the pool and the gather kernel finish together is a CANDIDATE, measured below.  The copy-engine path stays
        # available through `dma_fraction=` but is not chosen automatically.
        c0 = 0.15                                               # ms of launches / events per step on the host
        f2 = 0.0
        if t_host + c0 > t_gpu:
            # host stage (1 - f) t_host + c0 and zero-copy stage f t_zc run side by side (the gather kernel, 256 CTAs, overlaps
            # the step's kernels fully): balance the two
            f2 = min(1.0, max(0.0, (t_host + c0) / (t_host + t_zc)))
        # The model ignores what the two stages do to each other (the gather kernel shares the SMs and the host DRAM with
        # the step and with the pool), so when it asks for a zero-copy share the candidates are MEASURED on the real
        # pipeline -- a few steps each of: pool only, the modelled share, half of it -- and the fastest one is kept.
        trials = {}
        if f2 > 0.0:
            self.zc_fraction, self.dma_fraction = 0.0, 0.0
            for _ in range(6):                                  # both slots used, their CUDA graphs captured: not part of any trial
                self.submit(out_img, cloud, choose, idx)
            cands = {0.0, round(f2 / 2, 3), round(f2, 3)}
            if f2 >= 0.5:
                cands.add(1.0)                                  # cores so scarce that the kernel alone may win
            for f in sorted(cands):
                self.zc_fraction, self.dma_fraction = f, 0.0
                for _ in range(3):
                    self.submit(out_img, cloud, choose, idx)
                torch.cuda.synchronize(self.dev)
                t0 = time.perf_counter()
                for _ in range(trial_steps):
                    self.submit(out_img, cloud, choose, idx)
                torch.cuda.synchronize(self.dev)
                trials[f] = (time.perf_counter() - t0) / trial_steps * 1e3
            f2 = min(trials, key=trials.get)
            if trials[f2] > 0.97 * trials[0.0]:                 # within the noise of a few steps: keep the pool alone
                f2 = 0.0
        self.zc_fraction, self.dma_fraction = f2, 0.0
        self.calibration = dict(zero_copy_ms=t_zc, host_ms=t_host, dma_ms=t_dma, compute_ms=t_gpu, zero_copy_fraction=f2,
                                dma_fraction=0.0, host_fraction=1.0 - f2, measured_ms_per_step_by_zero_copy_fraction=trials)
        return self.calibration

    def _compute(self, s, B, gathered):
        """The step's kernels on the current stream (a CUDA-graph replay once the slot / batch size has been seen)."""
        src = self.d_emb[s][:B] if gathered else self.d_img[s][:B]

        def launch():
            ops.pose_pipeline(self.est, self.ref, src, self.d_cloud[s][:B], None if gathered else self.d_choose[s][:B], self.d_idx[s][:B],
                              iterations=self.iterations, canonical=self.canonical, out=self.d_pose[s][:B], gathered=gathered)
        key = (s, B, gathered)
        g = self.graphs.get(key)
        if g is None and self.use_graph and self.i >= 2:         # the first use of each slot runs (and warms up) eagerly
            try:
                cap = torch.cuda.Stream(device=self.dev)
                cap.wait_stream(torch.cuda.current_stream(self.dev))
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=cap):
                    launch()
                torch.cuda.current_stream(self.dev).wait_stream(cap)
            except Exception:                                     # capture not possible (e.g. profiling hooks on): plain launches
                g = False
            self.graphs[key] = g
        if g:
            g.replay()
        else:
            launch()

    def submit(self, out_img, cloud, choose, idx):
        """out_img [B,32,H,W] fp32: pinned host tensor (contiguous or channels_last), pageable host tensor or CUDA tensor;
        cloud [B,N,3] fp32, choose [B,1,N] int64, idx [B,1] int64 host (pinned for true async) or CUDA.  Returns immediately."""
        s = self.i % 2
        B = cloud.shape[0]
        main = torch.cuda.current_stream(self.dev)
        prof = self.cpu_ms
        t_ = time.perf_counter() if prof is not None else 0.0

        def lap(name):
            nonlocal t_
            if prof is not None:
                t1 = time.perf_counter(); prof[name] = prof.get(name, 0.0) + (t1 - t_) * 1e3; t_ = t1
        on_dev = out_img.is_cuda
        gathered = self.transfer == 'gather' and not on_dev
        k, kd, keep = B, 0, None                               # objects [0,k): zero-copy kernel, [k,k+kd): copy engine, [k+kd,B): host pool
        if gathered:
            nhwc = out_img.dim() == 4 and not out_img.is_contiguous() and out_img.is_contiguous(memory_format=torch.channels_last)
            if not out_img.is_pinned():
                k = 0                                          # pageable memory cannot be read by a kernel
            elif nhwc:
                k = B                                          # 128-byte lines: the kernel alone is PCIe-efficient
            else:
                if self.zc_fraction is None or self.dma_fraction is None:
                    self.calibrate(out_img, cloud, choose, idx)
                    s = self.i % 2                             # the calibration's trial steps went through the slots
                k = int(round(B * self.zc_fraction))
                kd = min(B - k, int(round(B * self.dma_fraction)))
            if kd > 0 and (self.d_img[s] is None or self.d_img[s].shape[0] < kd):
                self.d_img[s] = torch.empty((max(kd, 4), 32, self.crop_pixels), dtype=torch.float32, device=self.dev)
            if self.i >= 2:
                self.copied[s].synchronize()                  # the pinned staging buffer of slot s has been shipped
        def start_pool():
            ch = choose.reshape(B, -1)
            return ops.host_gather_begin(out_img, ch if ch.is_contiguous() else ch.contiguous(), self.h_stage[s], k + kd, B, self.host_threads)
        # Pool only: it starts first and the (short) enqueues below overlap it.  With a zero-copy / copy-engine share the
        # device work is enqueued first: once the pool threads run, this thread competes with them for the cores (8 ranks on
        # one host: 4 cores per rank) and the gather kernel would start late.
        pool_first = gathered and k + kd < B and k + kd == 0
        if pool_first:
            keep = start_pool()
        if self.i >= 2:
            self.copy_stream.wait_event(self.done[s])         # device slot s is free again
        with torch.cuda.stream(self.copy_stream):
            self.d_cloud[s][:B].copy_(cloud, non_blocking=True)
            self.d_choose[s][:B].copy_(choose.reshape(B, -1), non_blocking=True)
            self.d_idx[s][:B].copy_(idx.reshape(B), non_blocking=True)
            if gathered:
                if kd > 0:                                     # whole maps of these objects by DMA, columns picked on the device
                    self.d_img[s][:kd].copy_(out_img[k:k + kd].reshape(kd, 32, -1), non_blocking=True)
                if k > 0:
                    ops.gather_emb(out_img[:k], self.d_choose[s][:k], out=self.d_emb[s][:k])
                if kd > 0:
                    ops.gather_emb(self.d_img[s][:kd], self.d_choose[s][k:k + kd], out=self.d_emb[s][k:k + kd])
            elif not on_dev:
                self.d_img[s][:B].copy_(out_img.reshape(B, 32, -1), non_blocking=True)
        if gathered and k + kd < B and not pool_first:
            keep = start_pool()
        lap('enqueue_h2d')
        if gathered and k + kd < B:
            ops.host_gather_wait()
            lap('host_gather_wait')
            del keep
            with torch.cuda.stream(self.copy_stream):
                self.d_emb[s][k + kd:B].copy_(self.h_stage[s][k + kd:B], non_blocking=True)
        self.copied[s].record(self.copy_stream)
        main.wait_event(self.copied[s])
        lap('events')
        if on_dev:                                             # device-resident map: gather inside the front-end kernel
            ops.pose_pipeline(self.est, self.ref, out_img, self.d_cloud[s][:B], self.d_choose[s][:B], self.d_idx[s][:B],
                              iterations=self.iterations, canonical=self.canonical, out=self.d_pose[s][:B])
        else:
            self._compute(s, B, gathered)
        lap('compute_launch')
        self.h_pose[s][:B].copy_(self.d_pose[s][:B], non_blocking=True)
        self.done[s].record(main)
        lap('d2h')
        self.last = (s, B)
        self.i += 1

    def drain(self):
        """Wait for everything submitted; returns the poses [B,7] fp64 (wxyz, t) of the last batch (host)."""
        torch.cuda.synchronize(self.dev)
        if self.last is None:
            return None
        s, B = self.last
        return self.h_pose[s][:B].clone()


def estimate_poses(estimator, refiner, out_img, cloud, choose, idx, iterations=2, canonical=True, transfer='gather'):
    """One-shot convenience wrapper: numpy / CPU tensors in, numpy poses [B,7] (wxyz, t; metres) out."""
    t = [torch.as_tensor(np.ascontiguousarray(a)) if not isinstance(a, torch.Tensor) else a for a in (out_img, cloud, choose, idx)]
    B, N = t[1].shape[0], t[1].shape[1]
    r = Runner(estimator, refiner, B, N, int(np.prod(t[0].shape[2:])), iterations, canonical, transfer=transfer, use_graph=False)
    r.submit(*t)
    return r.drain().numpy()
