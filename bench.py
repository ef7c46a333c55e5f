#!/usr/bin/env python
"""bench.py -- pose frames/s (+ ICP registrations/s) of the B200-native pose-geometry hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[1], named in `config.workload`): DenseFusion PoseNet geometry +
2 canonical refine iterations, 500 sampled points / object, batch of 64 synthetic RGB-D frames
(one object per frame) PER GPU, colour-encoder output as input (the encoder stays on
PyTorch/cuDNN and is outside the graft).  A "step" = one pass of the hot path over
one batch.  One JSON line on stdout (rank 0 only).

  value      frames/s, inputs resident in HBM, CUDA events, max over ranks
  e2e        frames/s through the Python drop-in API with pinned HOST buffers: per step H2D of the
             step's inputs and D2H of the poses inside the timed region (double-buffered)
  roofline   the tcgen05 GEMM kernel (dominant): algorithmic FLOP/s vs the measured bf16 peak
  cpu_baseline  the oracle port (torch-CPU restatement of the reference) on the host cores, bounded sample
  icp        extra leg: back-projection + voxel grid + ICP registrations/s with their HBM roofline

--impl reference times the reference's CPU implementation of the same path (oracle port: the
reference is Python + open3d; the Python part is restated on torch-CPU, see oracle/densefusion.py).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH, NPTS, CROP, NUM_OBJ, REFINE_ITERS = 64, 500, (120, 160), 5, 2
WORKLOAD = 'C2: PoseNet geometry + %d canonical refine iters, %d pts/object, batch %d frames (1 object each) per GPU, ' \
           'encoder output [B,32,%d,%d] fp32 as input' % (REFINE_ITERS, NPTS, BATCH, CROP[0], CROP[1])

# reference-formulation FLOPs per point of the layers the tcgen05 GEMM kernel computes
GEMM_FLOPS_PER_PT = {
    'gemm.pn.conv2': 2 * 2 * 64 * 128, 'gemm.pn.conv5': 2 * 256 * 512, 'gemm.pn.conv6': 2 * 512 * 1024,
    'gemm.pn.heads1': 3 * 2 * 1408 * 640, 'gemm.pn.heads2': 3 * 2 * 640 * 256, 'gemm.pn.heads3': 3 * 2 * 256 * 128,
    'gemm.rf.conv2': 2 * 2 * 64 * 128, 'gemm.rf.conv5': 2 * 384 * 512, 'gemm.rf.conv6': 2 * 512 * 1024,
}
# executed MMA FLOPs per padded row and bf16 product: global-feature part of heads1 hoisted (K=384); multiplied by the
# number of split-bf16 products the layer runs (ape_net_get_passes: 3 for the per-point layers, 2 for the pooled ones)
GEMM_EXEC_PER_ROW_PASS = {
    'gemm.pn.conv2': 2 * 2 * 64 * 128, 'gemm.pn.conv5': 2 * 256 * 512, 'gemm.pn.conv6': 2 * 512 * 1024,
    'gemm.pn.heads1': 2 * 384 * 1920, 'gemm.pn.heads2': 3 * 2 * 640 * 256, 'gemm.pn.heads3': 3 * 2 * 256 * 128,
    'gemm.rf.conv2': 2 * 2 * 64 * 128, 'gemm.rf.conv5': 2 * 384 * 512, 'gemm.rf.conv6': 2 * 512 * 1024,
}
GEMM_LAYER_ORDER = ['conv2', 'conv5', 'conv6', 'heads1', 'heads2', 'heads3']


def gemm_exec_per_row(est, ref):
    pn, rf = est.get_passes(), ref.get_passes()
    out = {}
    for k, v in GEMM_EXEC_PER_ROW_PASS.items():
        net, layer = k.split('.')[1:]
        out[k] = v * bin((pn if net == 'pn' else rf)[GEMM_LAYER_ORDER.index(layer)]).count('1')
    return out


def measured_traffic(kernel, units=None):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full capture
    (profiles/traffic.json, written by tools/summarize_profiles.py); None when there is no capture."""
    p = os.path.join(ROOT, 'profiles', 'traffic.json')
    try:
        e = json.load(open(p))[kernel]
        b = e['dram_bytes_per_launch']
        if units is not None and e.get('registrations_per_launch'):      # captured at another launch size: scale per unit
            b = b / e['registrations_per_launch'] * units
        return b
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], bf16=d['bf16_tflops_sustained'], bf16_burst=d['bf16_tflops'], src='measured')
    return dict(hbm=6650.0, bf16=1400.0, bf16_burst=1590.0, src='fallback')


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md recipe)."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                       '-lms', '100'], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(',')]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), c[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            load = [s for s, p in zip(sm, pw) if p >= 0.5 * max(pw)] or sm
            out = dict(sm_mhz=statistics.median(load), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(pw))
        return out


def dist_env():
    rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    return rank, world, local


# ------------------------------------------------------------------------------------------- CPU arm
class CpuPoseArm:
    """The reference's CPU path for this workload, per sample as the reference runs it (batch 1 only, network.py:123):
    PoseNet geometry (cnn = Identity: the encoder output is the input, as for the B200 arm) + 2 canonical refine iterations
    + numpy pose composition.  kind 'reference': the reference's OWN modules (DenseFusion/lib/network.py, tools/utils.py,
    lib/transformations.py) compiled unmodified into oracle/_ref (oracle/ref_modules.py); kind 'port': the torch-CPU
    restatement oracle/densefusion.py when oracle/_ref was not built."""

    def __init__(self, threads=None, seed=7, n_inputs=8):
        import torch
        from oracle import ref_modules
        from autoposeestimation_b200 import synthetic as synth
        self.torch = torch
        self.threads = threads or os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        sd_e = synth.to_torch(synth.posenet_state_dict(seed, NUM_OBJ)); sd_r = synth.to_torch(synth.refiner_state_dict(seed + 1000, NUM_OBJ))
        self.inputs = [torch.from_numpy(a) for a in synth.posenet_inputs(seed, NPTS, CROP, NUM_OBJ, batch=n_inputs)]
        self.nb = n_inputs
        if ref_modules.available():
            self.kind = 'reference'
            self.mods = ref_modules.load()
            network = self.mods[0]
            self.est = network.PoseNet(NPTS, NUM_OBJ); self.est.cnn = torch.nn.Identity(); self.est.eval()
            self.ref = network.PoseRefineNet(NPTS, NUM_OBJ); self.ref.eval()
            self.est.load_state_dict(sd_e, strict=False); self.ref.load_state_dict(sd_r, strict=True)
            self.impl = 'reference modules compiled from /root/reference (oracle/_ref/DenseFusion/*.so), torch-CPU'
        else:
            self.kind = 'port'
            self.sd = (sd_e, sd_r)
            self.impl = 'torch-CPU oracle port (oracle/densefusion.py)'

    def one(self, b):
        t = self.inputs
        if self.kind == 'reference':
            from oracle import ref_modules
            return ref_modules.canonical_prediction(self.mods, self.est, self.ref, t[0][b:b + 1], t[1][b:b + 1], t[2][b:b + 1], t[3][b:b + 1],
                                                    NPTS, REFINE_ITERS)
        from oracle import densefusion as odf
        return odf.canonical_prediction(self.sd[0], self.sd[1], t[0][b:b + 1], t[1][b:b + 1], t[2][b:b + 1], t[3][b:b + 1], NUM_OBJ,
                                        iterations=REFINE_ITERS)

    def run(self, n_objects):
        """-> seconds for n_objects samples"""
        t0 = time.perf_counter()
        with self.torch.no_grad():
            for i in range(n_objects):
                self.one(i % self.nb)
        return time.perf_counter() - t0


def cpu_pose_frames_per_s(n_objects, threads=None, seed=7, min_seconds=0.0):
    """Bounded CPU sample beside the B200 line.  Returns (frames/s, seconds, threads, kind, impl)."""
    arm = CpuPoseArm(threads, seed)
    arm.run(1)                                                # warm-up
    done, dt = 0, 0.0
    while True:                                               # whole multiples of n_objects until min_seconds of CPU work are in
        dt += arm.run(n_objects); done += n_objects
        if dt >= min_seconds:
            break
    return done / dt, dt, arm.threads, arm.kind, arm.impl


def bench_config(n_sets=3, h2d=None):
    h2d = h2d if h2d is not None else BATCH * (32 * CROP[0] * CROP[1] * 4 + NPTS * 12 + NPTS * 8 + 8)
    return dict(workload=WORKLOAD, num_obj=NUM_OBJ, weights='random init (seeded), reference state_dict shapes',
                l2='no explicit flush: %d rotating input batches (%.0f MB each) and ~0.7 GB of activations per step exceed the 126 MB L2'
                   % (n_sets, h2d / 1e6),
                choose='host-sampled indices (given `choose`: bit-exact replay of the reference subset)',
                sharding='frames partitioned across ranks, no collective on the inference path')


def run_reference(args):
    """--impl reference: the same config, steps and warm-up as the B200 arm; a step = the 64 objects of one batch through the
    reference's per-sample CPU path on all host threads (~0.6 s per step on 16 cores)."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    warm = max(3, args.warmup)
    arm = CpuPoseArm()
    for _ in range(warm):
        arm.run(BATCH)
    t0 = time.perf_counter()
    secs = [arm.run(BATCH) for _ in range(args.steps)]
    total = time.perf_counter() - t0
    value = BATCH * args.steps / sum(secs)
    line = dict(metric='pose_frames_per_sec', value=value, unit='frames/s', n_gpus=args.gpus, steps=args.steps, warmup=warm,
                ms_per_step=1e3 * sum(secs) / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
                data='synthetic', impl='reference', config=bench_config(),
                cpu_baseline=dict(value=value, unit='frames/s', cores=arm.threads, kind=arm.kind,
                                  sample='%d steps x %d objects (whole batches), per-sample loop as the reference (batch 1 only, '
                                         'network.py:123), %s, %d threads' % (args.steps, BATCH, arm.impl, arm.threads)),
                e2e=dict(value=value, unit='frames/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0), wall_s=total)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- B200 arm
def profile_report(lib):
    n = lib.ape_profile_report(None, 0)
    buf = ctypes.create_string_buffer(n + 16)
    lib.ape_profile_report(buf, n + 16)
    out = {}
    for ln in buf.value.decode().splitlines():
        name, cnt, ms = ln.split()
        out[name] = (int(cnt), float(ms))
    return out


def icp_leg(torch, ops, lib, peaks, steps, rank=0, world=1, sync=None, reduce_max=None):
    """Extra leg (BASELINE configs 1/4), run by EVERY rank on its own frames (frames are independent: no collective on the
    data path; `sync` = barrier + device synchronize, `reduce_max` = max over ranks of a list of floats):
    (a) masked back-projection over 512 distinct 640x480 frames per rank (HBM roofline),
    (b) voxel grid + point-to-point ICP of 3552 registrations per rank built from 32 rendered frames.
    Whole-job rates = world * per-rank units / max-over-ranks device time."""
    from autoposeestimation_b200 import synthetic as synth
    sync = sync or torch.cuda.synchronize
    reduce_max = reduce_max or (lambda v: v)
    dev = torch.device('cuda', torch.cuda.current_device())
    F, H, W = 512, 480, 640                                   # 512 * 921 600 B = 472 MB > L2
    g = torch.Generator(device=dev).manual_seed(3 + rank)
    yy = torch.arange(H, device=dev).view(1, H, 1).float(); xx = torch.arange(W, device=dev).view(1, 1, W).float()
    cy = torch.randint(100, 380, (F, 1, 1), device=dev, generator=g).float(); cx = torch.randint(120, 520, (F, 1, 1), device=dev, generator=g).float()
    label = ((((yy - cy) / 45.0) ** 2 + ((xx - cx) / 60.0) ** 2) < 1.0).to(torch.uint8) * 255       # ~8.5 k px per frame
    depth = torch.randint(400, 900, (F, H, W), device=dev, generator=g, dtype=torch.int32)
    depth[torch.rand((F, H, W), device=dev, generator=g) < 0.02] = 0
    depth = depth.to(torch.int16)
    cam = torch.tensor([[320.0, 240.0, 615.0, 615.0]], dtype=torch.float64, device=dev).repeat(F, 1)
    r2c = torch.from_numpy(synth.hand_eye()).to(dev).repeat(F, 1, 1)
    cap = 12288
    pts, _, cnt = ops.surface_backproject(label, depth, cam, r2c, capacity=cap, want_pixels=False)     # warm-up
    torch.cuda.synchronize()
    lib.ape_profile_enable(1)
    for _ in range(min(steps, 50)):
        pts, _, cnt = ops.surface_backproject(label, depth, cam, r2c, capacity=cap, want_pixels=False)
    torch.cuda.synchronize()
    rep = profile_report(lib); lib.ape_profile_enable(0)
    ms_kernels = {k: rep[k][1] / rep[k][0] for k in ('surface_mask', 'surface_scan', 'surface_emit') if k in rep}
    # the launch as a whole (mask + scan + emit back to back, gaps included), CUDA events on the launching stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_bp = min(max(steps, 4), 50)
    sync()
    e0.record()
    for _ in range(n_bp):
        pts, _, cnt = ops.surface_backproject(label, depth, cam, r2c, capacity=cap, want_pixels=False)
    e1.record(); sync()
    ms_bp_rank = e0.elapsed_time(e1) / n_bp
    # (b) registrations: 3552 DISTINCT problems per rank -- 711 rendered frames of the 5-object scene (seeded camera poses on a
    # hemisphere), every (frame, object) surface cloud -> 2 mm voxel grid -> ICP against the object's perturbed 2000-point
    # model cloud.  3552 = 148 SMs x 24: whole rounds at 3, 4, 6 or 8 CTAs per SM.
    nreg, Lobj = 3552, 5
    scene = synth.Scene(3)
    nfr = (nreg + Lobj - 1) // Lobj
    poses = scene.camera_poses(500 + rank, nfr)
    lab_s = torch.empty((nfr, H, W), dtype=torch.uint8, device=dev); dep_s = torch.empty((nfr, H, W), dtype=torch.int16, device=dev)
    for c0 in range(0, nfr, 128):
        lab_s[c0:c0 + 128], dep_s[c0:c0 + 128] = scene.render(poses[c0:c0 + 128], seed=11 + 1000 * rank + c0, device=dev)
    cam_s = cam[:1].repeat(nfr, 1); r2c_s = torch.from_numpy(poses).to(dev)

    def once():
        o = ops.surface_backproject_multi(lab_s, dep_s, cam_s, r2c_s, [1, 2, 3, 4, 5], total_capacity=nfr * Lobj * 8192)
        vox, vc = ops.voxel_down_sample(o['points'], o['offsets'], 2.0, max_cloud_points=10240)
        return vox, o['offsets'][:nreg + 1].contiguous(), vc[:nreg].contiguous()

    src, so, vc = once()
    frames = [synth.render_ellipsoid_frame(100 + i) for i in range(4)]      # CPU arm below
    tgt = torch.from_numpy(np.concatenate([scene.models_pert[v % Lobj] for v in range(nreg)])).to(dev)
    to = torch.arange(0, nreg + 1, device=dev, dtype=torch.int32) * 2000
    T, info = ops.icp_p2p(src, so, tgt, to, 10.0, src_count=vc)             # warm-up
    n_icp = min(10, max(1, steps // 4))
    sync()
    e0.record()
    for _ in range(n_icp):
        T, info = ops.icp_p2p(src, so, tgt, to, 10.0, src_count=vc)
    e1.record(); sync()
    ms_icp_rank = e0.elapsed_time(e1) / n_icp
    vc_h = vc.cpu().numpy()
    ms_bp, ms_icp = reduce_max([ms_bp_rank, ms_icp_rank])    # max over ranks; every rank did the same amount of work
    nvalid = int(cnt.sum())
    bytes_bp = F * H * W * 3 + nvalid * 24                     # per rank
    iters = float(info[:, 2].mean()); ns_mean = float(np.mean(vc_h))
    bytes_icp = nreg * (24 * (ns_mean + 2000) + 128)
    # pair evaluations a brute-force NN would need (iters * Ns * Nt), the work the grid search avoids
    t0 = time.perf_counter(); once(); torch.cuda.synchronize(); prep_ms = (time.perf_counter() - t0) * 1e3
    converged = float((info[:, 0] > 0.9).double().mean())
    out = dict(backprojection=dict(
                   frames_per_s=world * F / ms_bp * 1e3, ms_per_launch=ms_bp, ms_per_kernel=ms_kernels, frames_per_launch=F, n_gpus=world,
                   kernels='surface_mask_kernel (stream label+depth -> validity bits) + surface_scan_kernel (slot prefix, list of non-empty '
                           '1024-pixel spans) + surface_emit_kernel (persistent warps, ordered fp64 points)',
                   roofline=dict(bound='hbm', achieved=bytes_bp / ms_bp / 1e6, peak=peaks['hbm'], unit='GB/s',
                                 frac=bytes_bp / ms_bp / 1e6 / peaks['hbm'], traffic=measured_traffic('surface_backproject'),
                                 peak_source=peaks['src'], algorithmic_bytes_per_launch=bytes_bp,
                                 note='per GPU; achieved = algorithmic bytes (921 600 B per frame + 24 B per valid pixel) / whole-call time')),
               icp=dict(registrations_per_s=world * nreg / ms_icp * 1e3, ms_per_launch=ms_icp, registrations_per_launch=nreg, n_gpus=world,
                        mean_source_points=ns_mean, target_points=2000, mean_iterations=iters,
                        mean_fitness=float(info[:, 0].mean()), mean_rmse_mm=float(info[:, 1].mean()), converged_fraction=converged,
                        problems='%d distinct (frame, object) registrations per rank (no replication)' % nreg,
                        roofline=dict(bound='hbm', achieved=bytes_icp / ms_icp / 1e6, peak=peaks['hbm'], unit='GB/s',
                                      frac=bytes_icp / ms_icp / 1e6 / peaks['hbm'], traffic=measured_traffic('icp_p2p', nreg), peak_source=peaks['src'],
                                      note='per GPU; compulsory bytes 24*(Ns+Nt)+128 per registration; the NN stage is FP64-ALU / latency '
                                           'bound (SURVEY 8d), so this fraction is expected to be small')),
               label_path_prep_ms=prep_ms)
    if world == 1:                                            # the reference's CPU path beside it (bounded sample, oracle port)
        out['cpu_baseline'] = cpu_label_path(frames[:4])
    return out


def reconstruct_leg(torch, ops, steps):
    """Extra leg: the SEQUENTIAL object-cloud reconstruction of main.py option 4 (create_pointcloud.py:232-317, "replicas
    only": each view registers to the cloud built so far): 30 views of one object -> get_surface for all views in one batched
    pass -> 29 x (voxel grids, one point-to-point ICP, transform, merge, voxel grid).  Latency, not throughput."""
    from autoposeestimation_b200 import synthetic as synth
    from autoposeestimation_b200.pc_reconstruction.create_pointcloud import get_surfaces_batch, reconstruct_run
    dev = torch.device('cuda', torch.cuda.current_device())
    n_views = 30
    scene = synth.Scene(5, n_objects=1)
    poses = scene.camera_poses(21, n_views)
    lab, dep = scene.render(poses, seed=3, device=dev, only_object=0)
    cam = torch.tensor([[synth.INTR['ppx'], synth.INTR['ppy'], synth.INTR['fx'], synth.INTR['fy']]], dtype=torch.float64, device=dev).repeat(n_views, 1)
    r2c = torch.from_numpy(poses).to(dev)

    def run():
        t0 = time.perf_counter()
        surfaces = get_surfaces_batch(lab, dep, cam, r2c, 20, 5.0, 20, 2.0)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        cloud = reconstruct_run(surfaces, 2.0, 10.0)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        return (t1 - t0) * 1e3, (t2 - t1) * 1e3, len(cloud), sum(len(s_) for s_ in surfaces) / n_views
    run()
    r = [run() for _ in range(3)]
    best = min(r, key=lambda x: x[0] + x[1])
    # the same loop driven view by view from Python (round-1 form: ~6 host round trips per view)
    surfaces = get_surfaces_batch(lab, dep, cam, r2c, 20, 5.0, 20, 2.0)
    host_ms = []
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        hc = reconstruct_run(surfaces, 2.0, 10.0, device_loop=False)
        torch.cuda.synchronize(); host_ms.append((time.perf_counter() - t0) * 1e3)
    return dict(views=n_views, surfaces_ms=best[0], register_and_merge_ms=best[1], ms_per_registration=best[1] / (n_views - 1),
                final_cloud_points=best[2], mean_surface_points=best[3],
                register_and_merge_ms_host_driven_loop=min(host_ms), final_cloud_points_host_driven_loop=len(hc),
                note='wall clock incl. host synchronisation (the loop is sequential by construction); surfaces = back-projection + voxel grid + '
                     'radius / statistical outlier filters of all 30 views in batched launches; register_and_merge = ape_reconstruct_run '
                     '(one call, sizes on the device, 512-thread ICP CTA per registration)')


def c4_leg(torch, ops, peaks, steps, rank=0, world=1, sync=None, reduce_max=None, lib=None):
    """Extra leg (BASELINE config 4 as written): batch pose-label generation over 10 000 synthetic frames x 5 objects, the
    frames sharded over the ranks (10 000 / world per rank, distinct frames: seeded camera poses on a hemisphere around a
    5-object scene), no collective.  Per chunk of frames, with NO host synchronisation: one-pass multi-label
    back-projection (921 600 B read once per FRAME, packed ragged clouds) -> 2 mm voxel grid -> point-to-point ICP of every
    (frame, object) cloud against the object's perturbed 2000-point model cloud (50 000 registrations per job)."""
    from autoposeestimation_b200 import synthetic as synth
    sync = sync or torch.cuda.synchronize
    reduce_max = reduce_max or (lambda v: v)
    dev = torch.device('cuda', torch.cuda.current_device())
    # chunk: 710 frames = 3550 registrations = 4 rounds of 6 resident ICP CTAs on each of the 148 SMs (the last chunk of a
    # rank is whatever is left)
    n_frames_job, L, chunk = 10000, 5, 710
    per_rank = n_frames_job // world
    scene = synth.Scene(3)
    poses = scene.camera_poses(1000 + rank, per_rank)
    H, W = 480, 640
    labels = torch.empty((per_rank, H, W), dtype=torch.uint8, device=dev)
    depths = torch.empty((per_rank, H, W), dtype=torch.int16, device=dev)
    for c0 in range(0, per_rank, 125):                          # input synthesis (untimed): analytic ray casting on the device
        lab, dep = scene.render(poses[c0:c0 + 125], seed=7 + 1000 * rank + c0, device=dev)
        labels[c0:c0 + 125] = lab; depths[c0:c0 + 125] = dep
    cam = torch.tensor([[synth.INTR['ppx'], synth.INTR['ppy'], synth.INTR['fx'], synth.INTR['fy']]], dtype=torch.float64, device=dev).repeat(per_rank, 1)
    r2c = torch.from_numpy(poses).to(dev)
    V = chunk * L
    tgt = torch.from_numpy(np.concatenate([scene.models_pert[v % L] for v in range(V)])).to(dev)
    to = torch.arange(0, V + 1, device=dev, dtype=torch.int32) * 2000
    cap_total = V * 10240
    results = []

    def run_chunk(c0):
        n = min(chunk, per_rank - c0)
        sl = slice(c0, c0 + n)
        out = ops.surface_backproject_multi(labels[sl], depths[sl], cam[sl], r2c[sl], [1, 2, 3, 4, 5], total_capacity=n * L * 10240)
        vox, vc = ops.voxel_down_sample(out['points'], out['offsets'], 2.0, max_cloud_points=10240)   # an object covers < 10 k pixels here
        T, info = ops.icp_p2p(vox, out['offsets'], tgt[:n * L * 2000], to[:n * L + 1], 10.0, src_count=vc)
        return out, vc, T, info

    starts = list(range(0, per_rank, chunk))
    n_chunks = len(starts)
    out, vc, T, info = run_chunk(0)                             # warm-up
    if starts[-1] != 0:
        run_chunk(starts[-1])
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for c0 in starts:
        out, vc, T, info = run_chunk(c0)
        results.append((out['offsets'][-1:], vc, info))
    e1.record(); sync()
    ms = reduce_max([e0.elapsed_time(e1)])[0]
    n_valid = int(sum(int(r[0]) for r in results))
    infos = torch.cat([r[2] for r in results])
    vcs = torch.cat([r[1] for r in results])
    frames = per_rank
    # back-projection alone on the same frames (multi-label, packed)
    e0.record()
    for c0 in starts:
        sl = slice(c0, min(c0 + chunk, per_rank))
        ops.surface_backproject_multi(labels[sl], depths[sl], cam[sl], r2c[sl], [1, 2, 3, 4, 5], total_capacity=cap_total)
    e1.record(); sync()
    kern = None
    if world == 1 and lib is not None:                          # where a chunk's time goes (CUDA events around every launch);
        lib.ape_profile_enable(1)                               # single process only: `sync` is a BARRIER under torchrun and
        run_chunk(0)                                            # nothing rank-dependent may call it
        torch.cuda.synchronize()
        kern = {k: v[1] for k, v in profile_report(lib).items()}
        lib.ape_profile_enable(0)
    ms_bp = reduce_max([e0.elapsed_time(e1)])[0]
    bytes_bp = frames * H * W * 3 + n_valid * 24
    return dict(frames_per_s=world * frames / ms * 1e3, registrations_per_s=world * frames * L / ms * 1e3, ms_total=ms,
                frames_per_rank=frames, objects_per_frame=L, registrations_per_rank=frames * L, n_gpus=world, chunk_frames=chunk,
                rank0_kernel_ms_first_chunk=kern,
                mean_source_points=float(vcs.float().mean()), mean_valid_pixels_per_view=n_valid / (frames * L),
                mean_iterations=float(infos[:, 2].mean()), mean_fitness=float(infos[:, 0].mean()), mean_rmse_mm=float(infos[:, 1].mean()),
                converged_fraction=float((infos[:, 0] > 0.9).double().mean()),
                backprojection=dict(frames_per_s=world * frames / ms_bp * 1e3, ms_total=ms_bp,
                                    roofline=dict(bound='hbm', achieved=bytes_bp / ms_bp / 1e6, peak=peaks['hbm'], unit='GB/s',
                                                  frac=bytes_bp / ms_bp / 1e6 / peaks['hbm'], peak_source=peaks['src'],
                                                  byte_model='921 600 B per FRAME (label + depth read once for all 5 objects) + 24 B per valid pixel written',
                                                  algorithmic_bytes=bytes_bp,
                                                  per_view_model_frac=(frames * L * H * W * 3 + n_valid * 24) / ms_bp / 1e6 / peaks['hbm'],
                                                  note='per_view_model_frac = the same time against 921 600 B per (frame, OBJECT), the byte model of the '
                                                       'single-label path (one scan per view): above 1 because the frame is read once for its 5 objects')),
                note='distinct frames, no replication; host synchronisation only at the end of the timed region')


def cpu_label_path(frames):
    """CPU arm of the label path on a bounded sample (oracle restatement: open3d 0.9 is not installable here): per frame the
    get_surface back-projection (vectorised numpy form of open3d_utils.py:172-192; the reference's literal per-pixel loop is
    timed on one frame), the 2 mm voxel grid and the point-to-point ICP against the 2000-point model (scipy cKDTree)."""
    from oracle import geometry as og, icp as oicp
    t_bp = t_icp = 0.0
    for fr in frames:
        t0 = time.perf_counter()
        pts, _ = og.surface_backproject(fr['label'], fr['depth'].astype(np.float64), fr['intr'], fr['robot2cam'])
        t1 = time.perf_counter()
        src = oicp.voxel_down_sample(pts, 2.0)
        oicp.registration_icp_p2p(src, fr['model'], 10.0)
        t2 = time.perf_counter()
        t_bp += t1 - t0; t_icp += t2 - t1
    fr = frames[0]
    t0 = time.perf_counter()
    og.surface_backproject_literal(fr['label'], fr['depth'].astype(np.float64), fr['intr'], fr['robot2cam'])
    t_lit = time.perf_counter() - t0
    return dict(kind='port', cores=1, sample='%d frames: numpy back-projection + voxel grid + ICP (oracle, one thread); literal per-pixel loop on 1 frame' % len(frames),
                backprojection_frames_per_s=len(frames) / t_bp, backprojection_literal_loop_frames_per_s=1.0 / t_lit,
                registrations_per_s=len(frames) / t_icp)


def adds_leg(torch, ops, peaks, steps, rank=0, world=1, sync=None, reduce_max=None):
    """Extra leg (BASELINE config 3): ADD / ADD-S evaluation with kNN for symmetric objects, 500 predicted points against a
    2 600-point model cloud, 100 k object instances sharded over 8 GPUs = 12 500 instances per rank (instances are
    independent: no collective; the final means are 2 scalars per rank).  Every instance has its own GT-posed target cloud
    (31 KB) and its own 500-point sample, as eval_linemod.py:118-130 sees them.  Timed twice: with the dataset's mix of
    symmetric classes (5 of 21) and with every instance symmetric (the kNN path for all)."""
    from autoposeestimation_b200 import synthetic as synth
    sync = sync or torch.cuda.synchronize
    reduce_max = reduce_max or (lambda v: v)
    dev = torch.device('cuda', torch.cuda.current_device())
    n_inst, n_model, n_pred = 12500, 2600, 500
    d = synth.adds_instances(2 + rank, n_inst, n_model_pts=n_model, n_pred_pts=n_pred)
    sub_np = d['subsample']; rest_np = np.setdiff1d(np.arange(n_model), sub_np)
    # the sampled 500 points first: row i of the GT-posed target is the partner of predicted point i (non-symmetric ADD)
    models = torch.from_numpy(np.concatenate([d['models'][:, sub_np], d['models'][:, rest_np]], axis=1)).to(dev)
    cls = torch.from_numpy(d['cls']).long().to(dev)
    q_gt = torch.from_numpy(d['q_gt']).to(dev); t_gt = torch.from_numpy(d['t_gt']).to(dev)
    w, x, y, z = q_gt.unbind(1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                     2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                     2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], 1).view(-1, 3, 3)
    target = torch.empty((n_inst, n_model, 3), dtype=torch.float32, device=dev)
    for c0 in range(0, n_inst, 2500):                          # setup only (untimed): GT-posed model cloud per instance
        sl = slice(c0, c0 + 2500)
        target[sl] = torch.bmm(models[cls[sl]], R[sl].transpose(1, 2)) + t_gt[sl, None, :]
    model_points = models[:, :n_pred][cls].contiguous()        # [B,500,3]
    q_pr = torch.from_numpy(d['q_pred']).to(dev); t_pr = torch.from_numpy(d['t_pred']).to(dev)
    sym_mix = torch.from_numpy(d['sym']).to(dev)[cls].contiguous()
    sym_all = torch.ones_like(sym_mix)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = max(3, min(steps // 10, 10))
    res = {}
    for name, sym in (('mixed', sym_mix), ('all_symmetric', sym_all)):
        dis = ops.add_metric(q_pr, t_pr, model_points, target, sym)     # warm-up
        sync()
        e0.record()
        for _ in range(n):
            dis = ops.add_metric(q_pr, t_pr, model_points, target, sym)
        e1.record(); sync()
        res[name] = (e0.elapsed_time(e1) / n, float(dis.double().mean()), float((dis < 0.02).double().mean()))
    ms_mix, ms_all = reduce_max([res['mixed'][0], res['all_symmetric'][0]])
    bytes_inst = (n_model + n_pred) * 12 + 2 * 28 + 4           # clouds + poses in, one distance out
    pair_flops = 8.0 * n_model * n_pred                          # brute-force pair evaluations x 8 FLOP
    fp32_peak = 148 * 128 * 2 * peaks.get('sm_max_mhz', 1965.0) * 1e6 / 1e12
    out = dict(instances_per_s=world * n_inst / ms_mix * 1e3, instances_per_s_all_symmetric=world * n_inst / ms_all * 1e3,
               ms_per_launch=ms_mix, ms_per_launch_all_symmetric=ms_all, instances_per_launch=n_inst, n_gpus=world,
               pred_points=n_pred, model_points=n_model, symmetric_fraction=float(sym_mix.float().mean()),
               mean_add_m=res['mixed'][1], below_2cm=res['mixed'][2],
               roofline=dict(bound='hbm', achieved=n_inst * bytes_inst / ms_mix / 1e6, peak=peaks['hbm'], unit='GB/s',
                             frac=n_inst * bytes_inst / ms_mix / 1e6 / peaks['hbm'], traffic=measured_traffic('add_metric'), peak_source=peaks['src'],
                             note='per GPU; compulsory bytes %d per instance; the symmetric (kNN) instances are FP32-ALU bound: all-symmetric run = '
                                  '%.1f TFLOP/s of brute-force pair arithmetic against ~%.0f TFLOP/s nominal fp32' %
                                  (bytes_inst, n_inst * pair_flops / ms_all / 1e9, fp32_peak)))
    if rank == 0:
        # the reference's own CUDA kNN (DenseFusion/lib/knn/src/cuda/knn.cu:217-263, compiled unmodified into oracle/_ref) as
        # "the kernel to beat" (SURVEY 8d iv) beside ape_knn on the same 500 x 2600 instances: checker code, timed only here
        try:
            from oracle import clib
            rlib = clib.ref_knn_cuda_lib()
            nb = 256
            refs = target[:nb].transpose(1, 2).contiguous()                  # [nb,3,2600]
            qrys = model_points[:nb].transpose(1, 2).contiguous()            # [nb,3,500]
            for _ in range(3):                                               # warm-up: the allocator's blocks for the index output exist
                idx_a = ops.knn(refs, qrys, 1, ops.KNN_ARITH_FMA)
            torch.cuda.synchronize()
            nk = 10
            e0.record()
            for _ in range(nk):
                ops.knn(refs, qrys, 1, ops.KNN_ARITH_FMA)
            e1.record(); torch.cuda.synchronize()
            ms_ours = e0.elapsed_time(e1) / nk
            cmp = dict(instances=nb, shape='500 queries x 2600 references, k=1', ape_knn_ms=ms_ours, ape_knn_instances_per_s=nb / ms_ours * 1e3)
            if rlib is not None:
                idx_r = torch.zeros_like(idx_a)
                scratch = torch.empty((n_model * n_pred,), dtype=torch.float32, device=dev)
                st = torch.cuda.current_stream().cuda_stream
                rlib.ref_knn_cuda(refs.data_ptr(), qrys.data_ptr(), idx_r.data_ptr(), scratch.data_ptr(), nb, 3, n_model, n_pred, 1, st)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(n):
                    rlib.ref_knn_cuda(refs.data_ptr(), qrys.data_ptr(), idx_r.data_ptr(), scratch.data_ptr(), nb, 3, n_model, n_pred, 1, st)
                e1.record(); torch.cuda.synchronize()
                ms_ref = e0.elapsed_time(e1) / n
                cmp.update(reference_knn_cu_ms=ms_ref, reference_knn_cu_instances_per_s=nb / ms_ref * 1e3, speedup=ms_ref / ms_ours,
                           indices_identical=bool(torch.equal(idx_a, idx_r)),
                           note='reference = its knn.cu recompiled for sm_100a, one instance per launch pair as knn.h:33-40 loops them, '
                                '10.4 MB distance matrix written and re-read per instance')
            out['knn_vs_reference_kernel'] = cmp
        except Exception as ex:
            out['knn_vs_reference_kernel'] = dict(error=repr(ex))
    if world == 1:
        from oracle import clib
        k = 512
        qs, ts_, mp, tg, sy = (a[:k].cpu().numpy() for a in (q_pr, t_pr, model_points, target, sym_mix))
        t0 = time.perf_counter(); done = 0
        while time.perf_counter() - t0 < 5.0:
            for i in range(k):
                clib.add_metric(qs[i], ts_[i], mp[i], tg[i], bool(sy[i]))
            done += k
        dt = time.perf_counter() - t0; k = done
        out['cpu_baseline'] = dict(kind='port', cores=1, sample='%d instances, C restatement of eval_linemod.py:118-130 + knn_cpu.cpp (one thread, %.1f s)' % (k, dt),
                                   instances_per_s=k / dt)
    return out


def live_leg(torch, ops, steps):
    """Extra leg: ONE live frame as main.py option 6 sees it (pipeline/utils.py:517-574): 5 detected objects x 1000 sampled
    points (the fork's num_points, :520), PoseNet + 2 canonical refine iterations, as launch-by-launch stream work and as
    one CUDA-graph replay (the whole block is graph-capturable: no host sync, no allocation inside ape_pose_pipeline)."""
    from autoposeestimation_b200 import _lib, synthetic as synth
    B, N = 5, 1000
    dev = torch.device('cuda', torch.cuda.current_device())
    est = ops.NetHandle(ops.NET_POSENET, synth.posenet_state_dict(7, NUM_OBJ), NUM_OBJ, B, N)
    ref = ops.NetHandle(ops.NET_REFINER, synth.refiner_state_dict(1007, NUM_OBJ), NUM_OBJ, B, N)
    d = [torch.from_numpy(a).to(dev) for a in synth.posenet_inputs(77, N, CROP, NUM_OBJ, batch=B)]
    d[0] = d[0].reshape(B, 32, -1).contiguous(); d[2] = d[2].reshape(B, N).contiguous(); d[3] = d[3].reshape(B).contiguous()
    poses = torch.empty((B, 7), dtype=torch.float64, device=dev)

    def run():
        ops.pose_pipeline(est, ref, d[0], d[1], d[2], d[3], iterations=REFINE_ITERS, canonical=True, out=poses)

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3                   # microseconds per frame

    for _ in range(5):
        run()
    n = max(20, min(steps, 200))
    us_stream = timed(run, n)
    ref_pose = poses.clone()
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        run()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        run()
    graph.replay(); torch.cuda.synchronize()
    same = bool(torch.equal(ref_pose, poses))
    us_graph = timed(graph.replay, n)
    lib = _lib.load()
    lib.ape_profile_enable(1)
    for _ in range(10):
        run()
    torch.cuda.synchronize()
    rep = profile_report(lib); lib.ape_profile_enable(0)
    kern = {}
    for k, v in rep.items():
        kk = 'gemm (tcgen05, all layers)' if k.startswith('gemm.') else k
        kern[kk] = kern.get(kk, 0.0) + v[1] / 10 * 1e3
    est.close(); ref.close()
    return dict(objects=B, points=N, refine_iters=REFINE_ITERS, us_per_frame_stream=us_stream, us_per_frame_cuda_graph=us_graph,
                kernel_us_per_frame=kern,
                frames_per_s_cuda_graph=1e6 / us_graph, graph_matches_stream_bitwise=same,
                note='device-resident inputs; one frame = 5 objects; latency mode of the same kernels as the headline')


def real_pipeline_leg(torch, steps, sd_e, sd_r):
    """Extra leg: the option-6 call as the reference makes it (pipeline/utils.py:556-571), batched: pinned RGB crops
    [64,3,120,160] + cloud + choose + idx on the HOST -> H2D -> colour encoder (PSPNet/ResNet-18 on PyTorch/cuDNN, outside
    the graft, default init) -> geometry kernels (device-resident map) -> poses on the host.  The encoder dominates."""
    from autoposeestimation_b200 import ops, synthetic as synth
    from autoposeestimation_b200.densefusion import network
    dev = torch.device('cuda', torch.cuda.current_device())
    torch.manual_seed(0)
    est = network.PoseNet(NPTS, NUM_OBJ); est.load_state_dict(synth.to_torch(sd_e), strict=False); est = est.to(dev).eval()
    ref = network.PoseRefineNet(NPTS, NUM_OBJ); ref.load_state_dict(synth.to_torch(sd_r), strict=False); ref = ref.to(dev).eval()
    _, cloud, choose, idx = synth.posenet_inputs(5, NPTS, CROP, NUM_OBJ, batch=BATCH)
    h = [torch.randn((BATCH, 3) + CROP).pin_memory()] + [torch.from_numpy(a).pin_memory() for a in (cloud, choose, idx)]
    h_pose = torch.empty((BATCH, 7), dtype=torch.float64).pin_memory()
    eh, rh = est._handle(BATCH, NPTS), ref._handle(BATCH, NPTS)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    res = {}
    for name, cl in (('nchw', False), ('channels_last', True)):
        if cl:
            est.cnn.to(memory_format=torch.channels_last)

        def step(timed=False):
            with torch.no_grad():
                d = [t.to(dev, non_blocking=True) for t in h]
                if cl:
                    d[0] = d[0].contiguous(memory_format=torch.channels_last)
                if timed:
                    ev[2].record()
                out_img = est.cnn(d[0])
                if timed:
                    ev[3].record()
                poses, _ = ops.pose_pipeline(eh, rh, out_img, d[1], d[2], d[3], iterations=REFINE_ITERS, canonical=True)
                h_pose.copy_(poses, non_blocking=True)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        n = max(3, min(steps, 20))
        ev[0].record()
        for _ in range(n):
            step()
        ev[1].record(); torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / n
        step(True); torch.cuda.synchronize()
        res[name] = dict(frames_per_s=BATCH / ms * 1e3, ms_per_step=ms, encoder_ms=ev[2].elapsed_time(ev[3]),
                         encoder_output_channels_last=bool(cl))
    return dict(res, h2d_bytes_per_step=sum(t.numel() * t.element_size() for t in h), d2h_bytes_per_step=BATCH * 7 * 8,
                note='host RGB crops -> cuDNN encoder (outside the graft) -> grafted geometry -> host poses, per batch of 64; with a '
                     'channels_last encoder the kernels gather 128-byte lines from its output')


TRAIN_BATCH, TRAIN_ITERS = 256, 2
# reference-formulation FLOPs of one refiner forward per object: 1 479 040 FLOP/pt + per-object heads
REFINER_FWD_FLOPS = 1479040 * NPTS + 2359296 + 2 * 128 * 7 * NUM_OBJ


def train_leg(torch, dist, lib, peaks, world, rank, dev, steps):
    """Extra leg (BASELINE config 5): PoseRefineNet training step, bf16, GLOBAL batch 256 objects x 500 points split over
    the ranks (strong scaling), 2 refinement iterations, NCCL all-reduce of the flat fp32 gradient, Adam.  Every rank runs
    it (the all-reduce is a collective); returns the dict on rank 0."""
    from autoposeestimation_b200 import synthetic as synth
    from autoposeestimation_b200.densefusion.train_refiner import RefinerTrainer
    B = TRAIN_BATCH // world
    g = torch.Generator(device=dev).manual_seed(4 + rank)
    pts = torch.randn((B, NPTS, 3), device=dev, generator=g) * 0.05
    emb = torch.randn((B, 32, NPTS), device=dev, generator=g)
    idx = torch.randint(0, NUM_OBJ, (B,), device=dev, generator=g)
    model = (torch.rand((B, NPTS, 3), device=dev, generator=g) - 0.5) * 0.2
    target = model + 0.01 * torch.randn((B, 1, 3), device=dev, generator=g)
    sd = synth.refiner_state_dict(1007, NUM_OBJ)
    sd['conv3_r.bias'] = sd['conv3_r.bias'].copy(); sd['conv3_r.bias'][0::4] += 1.0       # near-identity start, as a trained refiner
    trainer = RefinerTrainer(sd, NUM_OBJ, B, NPTS, sym_list=[0], iterations=TRAIN_ITERS)
    for _ in range(3):
        trainer.train_step(pts, emb, idx, target, model)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n = max(3, min(steps, 20))
    l0 = lib.ape_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        dis = trainer.train_step(pts, emb, idx, target, model)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches = int(lib.ape_launch_count() - l0)
    ms = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    ar_ms = 0.0
    ms_serial = None
    if world > 1:                                            # the same step with ONE all-reduce after the whole backward pass
        for _ in range(2):
            trainer.train_step(pts, emb, idx, target, model, overlap=False)
        dist.barrier(); torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(n):
            trainer.train_step(pts, emb, idx, target, model, overlap=False)
        s1.record(); dist.barrier(); torch.cuda.synchronize()
        t_s = torch.tensor([s0.elapsed_time(s1) / n], dtype=torch.float64, device=dev)
        dist.all_reduce(t_s, op=dist.ReduceOp.MAX)
        ms_serial = float(t_s)
    if world > 1:                                            # the collective on its own (flat fp32 gradient, NCCL sum)
        gbuf = trainer.h.grads
        for _ in range(3):
            dist.all_reduce(gbuf)
        torch.cuda.synchronize(); dist.barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(10):
            dist.all_reduce(gbuf)
        a1.record(); torch.cuda.synchronize()
        t_ar = torch.tensor([a0.elapsed_time(a1) / 10], dtype=torch.float64, device=dev)
        dist.all_reduce(t_ar, op=dist.ReduceOp.MAX)
        ar_ms = float(t_ar)
    out = None
    # per-kernel split: EVERY rank runs these steps (train_step contains the all-reduce); only rank 0 records events
    if rank == 0:
        lib.ape_profile_enable(1)
    for _ in range(3):
        trainer.train_step(pts, emb, idx, target, model)
    torch.cuda.synchronize()
    if rank == 0:
        rep = profile_report(lib); lib.ape_profile_enable(0)
        kern = {k: v[1] / 3 for k, v in rep.items()}
        flops = 3 * REFINER_FWD_FLOPS * TRAIN_ITERS * TRAIN_BATCH            # fwd + dgrad + wgrad, whole job
        gemm_ms = sum(v for k, v in kern.items() if k.startswith('gemm.'))
        out = dict(objects_per_s=TRAIN_BATCH / ms * 1e3, ms_per_step=ms, global_batch=TRAIN_BATCH, batch_per_gpu=B, points=NPTS,
                   iterations=TRAIN_ITERS, dtype='bf16 operands, fp32 accumulate / master weights / gradients',
                   allreduce_bytes=int(trainer.h.grads.numel() * 4) if world > 1 else 0, allreduce_ms=ar_ms,
                   ms_per_step_serial_allreduce=ms_serial,
                   allreduce='two blocks: conv6 + heads (89 %) on a side stream as soon as the last conv6 weight gradient is written, '
                             'conv1..conv5 after the backward pass' if world > 1 else 'none (one rank)',
                   gpu_launches_per_step=launches / n,
                   mean_dis=float(dis.mean()),
                   roofline=dict(bound='tensor', achieved=flops / ms / 1e9, peak=peaks['bf16'] * world, unit='TFLOP/s',
                                 frac=flops / ms / 1e9 / (peaks['bf16'] * world), algorithmic_flops_per_step=flops,
                                 note='reference-formulation FLOPs x3 (forward, dgrad, wgrad) over the whole step time'),
                   rank0_kernel_ms_per_step=kern, rank0_gemm_ms_per_step=gemm_ms)
    trainer.h.close()
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the B200 path has no CPU fallback; use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints its banner / INFO lines on STDOUT: keep stdout to the one JSON line by sending them to stderr
        # (the driver's rank check reads them there); NCCL_DEBUG itself is left as the caller set it
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from autoposeestimation_b200 import _lib, ops, synthetic as synth
    from autoposeestimation_b200.densefusion import estimate_poses          # public drop-in API (host buffers -> poses)
    lib = _lib.load()
    peaks = measured_peaks()
    dev = torch.device('cuda', local)

    sd_e = synth.posenet_state_dict(7, NUM_OBJ); sd_r = synth.refiner_state_dict(1007, NUM_OBJ)
    est = ops.NetHandle(ops.NET_POSENET, sd_e, NUM_OBJ, BATCH, NPTS)
    ref = ops.NetHandle(ops.NET_REFINER, sd_r, NUM_OBJ, BATCH, NPTS)
    n_sets = 3                                                # rotate distinct input batches (3 x 157 MB > L2)
    host_sets = [synth.posenet_inputs(1000 * rank + 10 + i, NPTS, CROP, NUM_OBJ, batch=BATCH) for i in range(n_sets)]
    dsets = [[torch.from_numpy(a).to(dev) for a in hs] for hs in host_sets]
    poses = torch.empty((BATCH, 7), dtype=torch.float64, device=dev)
    h2d = sum(a.nbytes for a in host_sets[0])                 # the whole host-resident step input (map + cloud + choose + idx)

    def step(i):
        d = dsets[i % n_sets]
        ops.pose_pipeline(est, ref, d[0], d[1], d[2], d[3], iterations=REFINE_ITERS, canonical=True, out=poses)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(3, args.warmup)):
        step(i)
    # ---- value: device-resident inputs
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    l0 = lib.ape_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    launches = int(lib.ape_launch_count() - l0)
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    value = world * BATCH * args.steps / ms_total * 1e3

    # ---- e2e: host pinned buffers -> public API -> host poses, hand-over inside the timed region
    pinned = [[torch.from_numpy(a).pin_memory() for a in hs] for hs in host_sets]
    small_bytes = sum(t.numel() * t.element_size() for t in pinned[0][1:])
    emb_bytes = BATCH * 32 * NPTS * 4
    d2h = BATCH * 7 * 8

    def e2e_run(transfer, sets, **kw):
        """-> (frames/s whole job, ms per step max over ranks, wall seconds, last host poses, runner)"""
        runner = estimate_poses.Runner(est, ref, BATCH, NPTS, CROP[0] * CROP[1], iterations=REFINE_ITERS, canonical=True,
                                       transfer=transfer, **kw)
        for i in range(max(4, args.warmup)):                  # both buffer slots warmed up, split calibrated, graphs captured
            runner.submit(*sets[i % len(sets)])
        runner.drain()
        barrier()
        t0 = time.perf_counter()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(args.steps):
            runner.submit(*sets[i % len(sets)])
        out = runner.drain()
        b.record()
        barrier()
        wall = time.perf_counter() - t0
        m = torch.tensor([max(a.elapsed_time(b), 0.0)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(m, op=dist.ReduceOp.MAX)
        return world * BATCH * args.steps / float(m) * 1e3, float(m) / args.steps, wall, out, runner

    threads = max(1, (os.cpu_count() or 1) // max(1, int(os.environ.get('LOCAL_WORLD_SIZE', world))))
    e2e_value, e2e_ms, wall_e2e, out_host, runner = e2e_run('gather', pinned, host_threads=threads)
    e2e_cal = runner.calibration
    # the same call on the two other hand-over paths (reported beside the headline): whole-map copy (round 1) and a
    # channels-last host map (zero-copy kernel alone)
    full_value, full_ms, _, out_full, _ = e2e_run('full', pinned[:2])
    pinned_cl = [[hs[0].contiguous(memory_format=torch.channels_last).pin_memory()] + hs[1:] for hs in pinned[:2]]
    cl_value, cl_ms, _, out_cl, _ = e2e_run('gather', pinned_cl)
    e2e_same = bool(torch.equal(out_full, out_cl))            # both ran sets[(steps-1) % 2] last
    del pinned_cl
    clocks = sampler.stop() if sampler else None
    train = None
    if not args.no_train:
        try:
            train = train_leg(torch, dist, lib, peaks, world, rank, dev, args.steps)
        except Exception as ex:                               # an extra leg must never take the headline down
            train = dict(error=repr(ex))
    # ---- label path (BASELINE configs 1/4): every rank works on its own frames, whole-job rates on rank 0
    label_leg = None
    if not args.no_icp:
        def reduce_max(vals):
            t = torch.tensor(vals, dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return [float(v) for v in t]
        if world == 1:
            try:
                label_leg = icp_leg(torch, ops, lib, peaks, args.steps, 0, 1, barrier, reduce_max)
            except Exception as ex:
                label_leg = dict(error=repr(ex))
        else:
            # no try/except here: a rank that dropped out of the leg would leave the others waiting in its collectives
            label_leg = icp_leg(torch, ops, lib, peaks, args.steps, rank, world, barrier, reduce_max)
        # ---- config 4 as written: 10 k frames x 5 objects sharded over the ranks
        if args.no_c4:
            pass
        elif world == 1:
            try:
                label_leg = dict(label_leg, label_c4=c4_leg(torch, ops, peaks, args.steps, 0, 1, barrier, reduce_max, lib))
            except Exception as ex:
                label_leg = dict(label_leg, label_c4=dict(error=repr(ex)))
        else:
            label_leg = dict(label_leg, label_c4=c4_leg(torch, ops, peaks, args.steps, rank, world, barrier, reduce_max, lib))
        torch.cuda.empty_cache()
        # ---- ADD / ADD-S evaluation (BASELINE config 3), same sharding rule
        if args.no_adds:
            pass
        elif world == 1:
            try:
                label_leg = dict(label_leg, add_metric=adds_leg(torch, ops, peaks, args.steps, 0, 1, barrier, reduce_max))
            except Exception as ex:
                label_leg = dict(label_leg, add_metric=dict(error=repr(ex)))
        else:
            label_leg = dict(label_leg, add_metric=adds_leg(torch, ops, peaks, args.steps, rank, world, barrier, reduce_max))

    line = None
    if rank == 0:
        # ---- roofline leg: same steps with per-launch CUDA events (after the timed regions, so they are unperturbed)
        lib.ape_profile_enable(1)
        psteps = min(args.steps, 50)
        for i in range(psteps):
            step(i)
        torch.cuda.synchronize()
        rep = profile_report(lib); lib.ape_profile_enable(0)
        rep = {k: (v[0] * args.steps / psteps, v[1] * args.steps / psteps) for k, v in rep.items()}   # normalise to args.steps
        gemm_ms = sum(v[1] for k, v in rep.items() if k.startswith('gemm.')) / args.steps
        all_ms = sum(v[1] for v in rep.values()) / args.steps
        refine = {k: (REFINE_ITERS if k.startswith('gemm.rf') else 1) for k in GEMM_FLOPS_PER_PT}
        # only the layers the GEMM kernel actually ran count (a layer fused into another kernel contributes neither FLOPs nor
        # time to this figure)
        in_gemm = [k for k in GEMM_FLOPS_PER_PT if k in rep or k in ('gemm.pn.heads1', 'gemm.pn.heads2') and 'gemm.pn.heads12' in rep]
        alg = sum(GEMM_FLOPS_PER_PT[k] * refine[k] for k in in_gemm) * BATCH * NPTS
        rows = BATCH * ((NPTS + 127) // 128 * 128)
        epr = gemm_exec_per_row(est, ref)
        exe = sum(epr[k] * refine[k] for k in in_gemm) * rows
        fpt = dict(GEMM_FLOPS_PER_PT)
        # the back-to-back heads kernel (gemm_tc4.cuh) does the work of two layers in one launch
        fpt['gemm.pn.heads12'] = fpt['gemm.pn.heads1'] + fpt['gemm.pn.heads2']
        epr['gemm.pn.heads12'] = epr['gemm.pn.heads1'] + epr['gemm.pn.heads2']
        layers = {k: dict(launches_per_step=rep[k][0] / args.steps, ms_per_step=rep[k][1] / args.steps,
                          algorithmic_tflops=fpt[k] * BATCH * NPTS * (rep[k][0] / args.steps) / (rep[k][1] / args.steps) / 1e9,
                          executed_bf16_tflops=epr[k] * rows * (rep[k][0] / args.steps) / (rep[k][1] / args.steps) / 1e9)
                  for k in fpt if k in rep}
        roofline = dict(bound='tensor', kernel='tcgen05 split-bf16 GEMM kernels (gemm_tc2 / gemm_tc4, %d launches/step)' % sum(
            round(rep[k][0] / args.steps) for k in rep if k.startswith('gemm.')),
            achieved=alg / gemm_ms / 1e9, peak=peaks['bf16'], unit='TFLOP/s', frac=alg / gemm_ms / 1e9 / peaks['bf16'],
            traffic=measured_traffic('gemm'), traffic_source='profiles/traffic.json (ncu --set full, bytes per GEMM launch)',
            peak_source=peaks['src'] + ' bf16_tflops_sustained',
            algorithmic_flops_per_step=alg, layers_in_gemm_kernel=sorted(in_gemm),
            layers_fused_elsewhere=sorted(k for k in GEMM_FLOPS_PER_PT if k not in in_gemm),
            executed_bf16_tflops=exe / gemm_ms / 1e9,
            executed_frac=exe / gemm_ms / 1e9 / peaks['bf16'], gemm_ms_per_step=gemm_ms, all_kernels_ms_per_step=all_ms,
            gemm_share_of_kernel_time=gemm_ms / all_ms, layers=layers,
            other_kernels_ms_per_step={k: v[1] / args.steps for k, v in rep.items() if not k.startswith('gemm.')},
            split_bf16_products=dict(posenet=est.get_passes(), refiner=ref.get_passes()[:3], layers=GEMM_LAYER_ORDER,
                                     key='7 = A_lo*W_hi + A_hi*W_lo + A_hi*W_hi, 6 = without A_lo*W_hi (layers whose output is pooled over the points)'),
            note='achieved = reference-formulation FLOPs (SURVEY 8d: 7.94 MFLOP/pt PoseNet + 1.48 MFLOP/pt per refine iter, GEMM layers) / '
                 'summed GEMM kernel time; executed = the split-bf16 products actually run (3 per-point layers, 2 pooled layers) on padded rows '
                 'with the global feature hoisted')
        # ---- CPU baseline beside it (bounded sample)
        # rank 0 at N=1 only (under torchrun OMP_NUM_THREADS=1 would make it a 1-thread number that looks like a regression)
        cpu_fps, cpu_s, cpu_threads, cpu_kind, cpu_impl = cpu_pose_frames_per_s(32, min_seconds=10.0) if world == 1 else (None, 0.0, 0, None, None)
        cpu_n = int(round(cpu_fps * cpu_s)) if cpu_fps else 0
        extra = label_leg
        if train is not None:
            extra = dict(extra or {}, refiner_training=train)
        try:
            extra = dict(extra or {}, live_frame=live_leg(torch, ops, args.steps))
        except Exception as ex:
            extra = dict(extra or {}, live_frame=dict(error=repr(ex)))
        if not args.no_icp:
            try:
                extra = dict(extra or {}, reconstruct_30_views=reconstruct_leg(torch, ops, args.steps))
            except Exception as ex:
                extra = dict(extra or {}, reconstruct_30_views=dict(error=repr(ex)))
        if world == 1:
            try:
                extra = dict(extra or {}, real_pipeline=real_pipeline_leg(torch, args.steps, sd_e, sd_r))
            except Exception as ex:
                extra = dict(extra or {}, real_pipeline=dict(error=repr(ex)))
        line = dict(metric='pose_frames_per_sec', value=value, unit='frames/s', n_gpus=world, steps=args.steps, warmup=max(3, args.warmup),
                    ms_per_step=ms_total / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='bf16x3',
                    data='synthetic', impl='b200',
                    config=bench_config(n_sets, h2d),
                    e2e=dict(value=e2e_value, unit='frames/s', h2d_bytes_per_step=emb_bytes + small_bytes, d2h_bytes_per_step=d2h,
                             ms_per_step=e2e_ms, wall_s=wall_e2e, frac_of_value=e2e_value / value,
                             h2d_gbs_per_gpu=(emb_bytes + small_bytes) / e2e_ms / 1e6,
                             host_map_bytes_per_step=h2d, transfer='gather', host_threads=threads, calibration=e2e_cal,
                             note='host-resident fp32 encoder maps [64,32,120,160] (pinned, the reference encoder\'s NCHW layout); only the 500 '
                                  'sampled columns per channel cross PCIe: the batch is split between the zero-copy gather kernel (reads the pinned '
                                  'map in place; 32-byte PCIe reads, ~8x the useful bytes on the wire) and the host thread pool (gather into pinned '
                                  'staging + one H2D); cloud / choose / idx H2D and poses D2H inside the timed region; kernels replayed from a CUDA graph',
                             api='autoposeestimation_b200.densefusion.estimate_poses.Runner (pinned host buffers, hand-over / compute double-buffered)',
                             full_map=dict(value=full_value, ms_per_step=full_ms, h2d_bytes_per_step=h2d, h2d_gbs_per_gpu=h2d / full_ms / 1e6,
                                           note='round-1 path: cudaMemcpyAsync of the whole map, gather in the front-end kernel'),
                             channels_last=dict(value=cl_value, ms_per_step=cl_ms, h2d_bytes_per_step=emb_bytes + small_bytes,
                                                note='same maps in torch.channels_last memory format: one sampled point = one 128-byte line, '
                                                     'zero-copy gather kernel alone', poses_equal_full_map_path=e2e_same)),
                    gpu_launches=launches, clocks=clocks, roofline=roofline,
                    cpu_baseline=(dict(value=cpu_fps, unit='frames/s', cores=cpu_threads, kind=cpu_kind,
                                       sample='%d objects of the same workload, per-sample loop, %s (%.1f s)' % (cpu_n, cpu_impl, cpu_s))
                                  if cpu_fps is not None else None),
                    extra=extra, checksum=float(out_host.sum()))
        line['summary'] = make_summary(line)                   # LAST key: every BASELINE metric in the tail the driver keeps
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def make_summary(line):
    """Compact per-N digest of every BASELINE metric (pose frames/s, ICP registrations/s, roofline fractions, ADD-S, training)."""
    def g(d, *ks):
        for k in ks:
            d = d.get(k) if isinstance(d, dict) else None
        return round(d, 4) if isinstance(d, float) else d
    x = line.get('extra') or {}
    return dict(n=line['n_gpus'], fps=round(line['value']), ms=round(line['ms_per_step'], 4), e2e_fps=round(line['e2e']['value']),
                e2e_frac=g(line, 'e2e', 'frac_of_value'), e2e_h2d_gbs=g(line, 'e2e', 'h2d_gbs_per_gpu'),
                e2e_cl_fps=g(line, 'e2e', 'channels_last', 'value'), e2e_full_fps=g(line, 'e2e', 'full_map', 'value'),
                gemm_frac=g(line, 'roofline', 'frac'), gemm_exec_frac=g(line, 'roofline', 'executed_frac'),
                bp_fps=g(x, 'backprojection', 'frames_per_s'), bp_gbs=g(x, 'backprojection', 'roofline', 'achieved'),
                bp_frac=g(x, 'backprojection', 'roofline', 'frac'),
                icp_rps=g(x, 'icp', 'registrations_per_s'), icp_hbm_frac=g(x, 'icp', 'roofline', 'frac'),
                c4_fps=g(x, 'label_c4', 'frames_per_s'), c4_rps=g(x, 'label_c4', 'registrations_per_s'),
                c4_bp_frac=g(x, 'label_c4', 'backprojection', 'roofline', 'frac'),
                adds_ips=g(x, 'add_metric', 'instances_per_s'), adds_sym_ips=g(x, 'add_metric', 'instances_per_s_all_symmetric'),
                knn_vs_ref=g(x, 'add_metric', 'knn_vs_reference_kernel', 'speedup'),
                train_ms=g(x, 'refiner_training', 'ms_per_step'), train_ops=g(x, 'refiner_training', 'objects_per_s'),
                train_ar_ms=g(x, 'refiner_training', 'allreduce_ms'), train_frac=g(x, 'refiner_training', 'roofline', 'frac'),
                live_us=g(x, 'live_frame', 'us_per_frame_cuda_graph'), recon30_ms=g(x, 'reconstruct_30_views', 'register_and_merge_ms'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-icp', action='store_true', help='skip the extra ICP / back-projection leg')
    ap.add_argument('--no-train', action='store_true', help='skip the extra refiner-training leg (BASELINE config 5)')
    ap.add_argument('--no-c4', action='store_true', help='skip the 10 000-frame label-generation leg (BASELINE config 4; profiling runs)')
    ap.add_argument('--no-adds', action='store_true', help='skip the ADD / ADD-S evaluation leg (BASELINE config 3; profiling runs)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
