"""Tensor-level wrappers over the C-ABI (one function per entry point of include/ape_b200.h).

Inputs are CUDA torch tensors (torch is only the allocator / stream provider here); every
function launches on torch's current stream and returns device tensors.  There is no CPU path.
"""
import torch

from . import _lib
from ._lib import check, ptr, require_cuda, stream_ptr

import functools

KNN_ARITH_CPU = 0
KNN_ARITH_FMA = 1


def _on_tensor_device(fn):
    """Run `fn` with the CUDA device of its first CUDA-tensor argument (or of `self.device`) current, so that allocations,
    the stream handed to the library and the kernels all belong to the device the data lives on, whatever device is
    current in the caller (the reference's modules follow their parameters' device the same way)."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = None
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.Tensor) and a.is_cuda:
                dev = a.device
                break
            d = getattr(a, 'device', None)
            if dev is None and isinstance(d, torch.device) and d.type == 'cuda' and not isinstance(a, torch.Tensor):
                dev = d
                break
        if dev is None or dev.index is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapper


def _c(t, dtype):
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


@_on_tensor_device
def backproject_choose(depth, bbox, choose, cam, frame_of=None):
    """a3.  depth [F,H,W] uint16 (torch.int16/uint16 storage accepted), bbox [B,4] int32,
    choose [B,N] int64, cam [B,5] fp32 (ppx,ppy,fx,fy,depth_scale) -> cloud [B,N,3] fp32."""
    require_cuda(depth, bbox, choose, cam, frame_of)
    assert depth.dtype in (torch.uint16, torch.int16) and depth.dim() == 3
    depth = depth.contiguous(); bbox = _c(bbox, torch.int32); choose = _c(choose, torch.int64); cam = _c(cam, torch.float32)
    if frame_of is not None:
        frame_of = _c(frame_of, torch.int32)
    B, N = choose.shape
    F, H, W = depth.shape
    cloud = torch.empty((B, N, 3), dtype=torch.float32, device=depth.device)
    check(_lib.load().ape_backproject_choose(ptr(depth), F, H, W, ptr(frame_of), ptr(bbox), ptr(choose), ptr(cam),
                                             B, N, ptr(cloud), stream_ptr()), 'ape_backproject_choose')
    return cloud


@_on_tensor_device
def mask_bbox_choose(label, depth, cam, n_points, frame_of=None, label_value=None, seeds=None, want_cloud=True):
    """f-2 (a1+a2+a3 on the device).  label [F,H,W] uint8, depth [F,H,W] uint16/int16 storage, cam [B,5] fp32 ->
    dict(bbox [B,4] int32, n_candidates [B] int32, choose [B,N] int64, cloud [B,N,3] fp32 or None)."""
    require_cuda(label, depth, cam, frame_of, label_value, seeds)
    assert label.dtype == torch.uint8 and depth.dtype in (torch.uint16, torch.int16)
    label = label.contiguous(); depth = depth.contiguous(); cam = _c(cam, torch.float32)
    F, H, W = label.shape
    B = cam.shape[0]
    dev = label.device
    if frame_of is not None:
        frame_of = _c(frame_of, torch.int32)
    if label_value is not None:
        label_value = _c(label_value, torch.uint8)
    if seeds is not None:
        if seeds.dtype != torch.int32:                           # any integer dtype: keep the low 32 bits (same bit pattern)
            v = seeds.to(torch.int64) & 0xffffffff
            seeds = torch.where(v >= 2 ** 31, v - 2 ** 32, v).to(torch.int32)
        seeds = seeds.contiguous()
    bbox = torch.zeros((B, 4), dtype=torch.int32, device=dev)
    ncand = torch.zeros((B,), dtype=torch.int32, device=dev)
    choose = torch.zeros((B, n_points), dtype=torch.int64, device=dev)
    cloud = torch.zeros((B, n_points, 3), dtype=torch.float32, device=dev) if want_cloud else None
    check(_lib.load().ape_mask_bbox_choose(ptr(label), ptr(depth), F, H, W, ptr(frame_of), ptr(label_value), ptr(seeds), ptr(cam),
                                           B, int(n_points), ptr(bbox), ptr(ncand), ptr(choose), ptr(cloud), stream_ptr()),
          'ape_mask_bbox_choose')
    return dict(bbox=bbox, n_candidates=ncand, choose=choose, cloud=cloud)


@_on_tensor_device
def surface_backproject(label, depth, cam, robot2cam, capacity, frame_of=None, label_value=None, want_pixels=True):
    """a4.  label [F,H,W] uint8, depth [F,H,W] uint16, cam [V,4] fp64, robot2cam [V,4,4] fp64.
    Returns (points [V,capacity,3] fp64, pixel_index [V,capacity] int32 or None, counts [V] int32)."""
    require_cuda(label, depth, cam, robot2cam)
    assert label.dtype == torch.uint8 and depth.dtype in (torch.uint16, torch.int16)
    label = label.contiguous(); depth = depth.contiguous()
    cam = _c(cam, torch.float64); robot2cam = _c(robot2cam, torch.float64)
    F, H, W = label.shape
    V = cam.shape[0]
    if frame_of is not None:
        frame_of = _c(frame_of, torch.int32)
    if label_value is not None:
        label_value = _c(label_value, torch.uint8)
    dev = label.device
    points = torch.empty((V, capacity, 3), dtype=torch.float64, device=dev)
    pix = torch.empty((V, capacity), dtype=torch.int32, device=dev) if want_pixels else None
    counts = torch.empty((V,), dtype=torch.int32, device=dev)
    lib = _lib.load()
    work = torch.empty(((int(lib.ape_surface_work_bytes(V, H, W)) + 3) // 4,), dtype=torch.int32, device=dev)
    check(lib.ape_surface_backproject(ptr(label), ptr(depth), F, H, W, ptr(frame_of), ptr(label_value), ptr(cam),
                                      ptr(robot2cam), V, capacity, ptr(points), ptr(pix), ptr(counts), ptr(work),
                                      stream_ptr()), 'ape_surface_backproject')
    return points, pix, counts


@_on_tensor_device
def surface_backproject_multi(label, depth, cam, robot2cam, label_values, total_capacity, want_pixels=False):
    """a4 for frames carrying several object labels (config 4), one pass per frame, PACKED output, no host sync.
    label [F,H,W] uint8, depth [F,H,W] uint16/int16 storage, cam [F,4] fp64, robot2cam [F,4,4] fp64 (per frame), label_values:
    list of 1..8 non-zero label values -> dict(points [total_capacity,3] fp64, offsets [F*L+1] int32, counts [F*L] int32,
    pixel_index [total_capacity] int32 or None): view v = f * L + l occupies points[offsets[v] : offsets[v+1])."""
    import ctypes
    require_cuda(label, depth, cam, robot2cam)
    assert label.dtype == torch.uint8 and depth.dtype in (torch.uint16, torch.int16)
    label = label.contiguous(); depth = depth.contiguous()
    cam = _c(cam, torch.float64); robot2cam = _c(robot2cam, torch.float64)
    F, H, W = label.shape
    L = len(label_values)
    V = F * L
    dev = label.device
    points = torch.empty((int(total_capacity), 3), dtype=torch.float64, device=dev)
    pix = torch.empty((int(total_capacity),), dtype=torch.int32, device=dev) if want_pixels else None
    counts = torch.empty((V,), dtype=torch.int32, device=dev)
    offsets = torch.empty((V + 1,), dtype=torch.int32, device=dev)
    lib = _lib.load()
    work = torch.empty(((int(lib.ape_surface_work_bytes(V, H, W)) + 3) // 4,), dtype=torch.int32, device=dev)
    vals = (ctypes.c_uint8 * L)(*[int(v) for v in label_values])
    check(lib.ape_surface_backproject_multi(ptr(label), ptr(depth), F, H, W, vals, L, ptr(cam), ptr(robot2cam), int(total_capacity),
                                            ptr(points), ptr(pix), ptr(counts), ptr(offsets), ptr(work), stream_ptr()),
          'ape_surface_backproject_multi')
    return dict(points=points, offsets=offsets, counts=counts, pixel_index=pix)


@_on_tensor_device
def knn(ref, query, k=1, arith=KNN_ARITH_CPU, out=None):
    """a14.  ref [B,D,N], query [B,D,M] fp32 -> idx [B,k,M] int64, 1-based."""
    require_cuda(ref, query)
    ref = _c(ref, torch.float32); query = _c(query, torch.float32)
    B, D, N = ref.shape
    M = query.shape[2]
    if out is None:
        out = torch.empty((B, k, M), dtype=torch.int64, device=ref.device)
    check(_lib.load().ape_knn(ptr(ref), ptr(query), ptr(out), B, D, N, M, k, arith, stream_ptr()), 'ape_knn')
    return out


@_on_tensor_device
def add_metric(quat, trans, model_points, target, symmetric, want_index=False):
    """a13/a15.  quat [B,4], trans [B,3]; model_points [B,Mq,3] or [Mq,3] (shared); target [B,Nt,3] or
    [Nt,3]; symmetric [B] uint8/bool.  Returns dis [B] fp32 (and nn_index [B,Mq] int32)."""
    require_cuda(quat, trans, model_points, target)
    quat = _c(quat, torch.float32); trans = _c(trans, torch.float32)
    model_points = _c(model_points, torch.float32); target = _c(target, torch.float32)
    B = quat.shape[0]
    mq, nt = model_points.shape[-2], target.shape[-2]
    ms = 0 if model_points.dim() == 2 else mq * 3
    ts = 0 if target.dim() == 2 else nt * 3
    sym = _c(symmetric.to(torch.uint8), torch.uint8)
    dis = torch.empty((B,), dtype=torch.float32, device=quat.device)
    nn = torch.empty((B, mq), dtype=torch.int32, device=quat.device) if want_index else None
    check(_lib.load().ape_add_metric(ptr(quat), ptr(trans), ptr(model_points), ms, mq, ptr(target), ts, nt, ptr(sym), B,
                                     ptr(dis), ptr(nn), stream_ptr()), 'ape_add_metric')
    return (dis, nn) if want_index else dis


@_on_tensor_device
def add_metric_std(quat, trans, model_points, target, symmetric):
    """f-3.  As add_metric, plus the unbiased std of the per-point distances: -> (dis [B], std [B]).  With 2-D
    model_points / target ([M,3]) they are shared by all B poses: the per-point candidate poses of `Loss` (loss.py:30-50)."""
    require_cuda(quat, trans, model_points, target)
    quat = _c(quat, torch.float32); trans = _c(trans, torch.float32)
    model_points = _c(model_points, torch.float32); target = _c(target, torch.float32)
    B = quat.shape[0]
    mq, nt = model_points.shape[-2], target.shape[-2]
    ms = 0 if model_points.dim() == 2 else mq * 3
    ts = 0 if target.dim() == 2 else nt * 3
    sym = _c(symmetric.to(torch.uint8), torch.uint8)
    dis = torch.empty((B,), dtype=torch.float32, device=quat.device)
    std = torch.empty((B,), dtype=torch.float32, device=quat.device)
    check(_lib.load().ape_add_metric_std(ptr(quat), ptr(trans), ptr(model_points), ms, mq, ptr(target), ts, nt, ptr(sym), B,
                                         ptr(dis), ptr(std), stream_ptr()), 'ape_add_metric_std')
    return dis, std


@_on_tensor_device
def estimator_loss(pred_r, pred_t, pred_c, points, model_points, target, symmetric, w, want_grad=True, want_pred=True):
    """a12.  `Loss` (lib/loss.py:12-73) forward + backward for the N candidate poses of one object.  pred_r [N,4], pred_t [N,3],
    pred_c [N], points [N,3], model_points / target [M,3] fp32 CUDA ->
    dict(loss [] , dis_best [], which_max [1] int32, dis [N], std [N], d_r [N,4], d_t [N,3], d_c [N] (d loss / d inputs, or None),
    new_points [1,N,3], new_target [1,M,3], pred [N,M,3] or None)."""
    require_cuda(pred_r, pred_t, pred_c, points, model_points, target)
    pred_r = _c(pred_r, torch.float32).reshape(-1, 4); N = pred_r.shape[0]
    pred_t = _c(pred_t, torch.float32).reshape(N, 3); pred_c = _c(pred_c, torch.float32).reshape(N)
    points = _c(points, torch.float32).reshape(N, 3)
    model_points = _c(model_points, torch.float32).reshape(-1, 3); M = model_points.shape[0]
    target = _c(target, torch.float32).reshape(M, 3)
    dev = pred_r.device
    f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    loss_dis, which = f(2), torch.empty((1,), dtype=torch.int32, device=dev)
    dis, std, term = f(N), f(N), f(N)
    d_r, d_t, d_c = (f(N, 4), f(N, 3), f(N)) if want_grad else (None, None, None)
    newp, newt = f(1, N, 3), f(1, M, 3)
    pred = f(N, M, 3) if want_pred else None
    check(_lib.load().ape_estimator_loss(ptr(pred_r), ptr(pred_t), ptr(pred_c), ptr(points), ptr(model_points), ptr(target), N, M,
                                         int(bool(symmetric)), float(w), ptr(loss_dis), ptr(which), ptr(dis), ptr(std), ptr(term),
                                         ptr(d_r), ptr(d_t), ptr(d_c), ptr(newp), ptr(newt), ptr(pred), stream_ptr()),
          'ape_estimator_loss')
    return dict(loss=loss_dis[0], dis_best=loss_dis[1], which_max=which, dis=dis, std=std, d_r=d_r, d_t=d_t, d_c=d_c,
                new_points=newp, new_target=newt, pred=pred)


@_on_tensor_device
def icp_p2p(source, src_offset, target, tgt_offset, threshold, rel_fitness=1e-2, rel_rmse=1e-2, max_iter=100, init=None, src_count=None):
    """a5.  Ragged batch: source [S,3] fp64 with src_offset [R+1] int32, target [T,3] fp64 with tgt_offset [R+1].
    src_count [R] int32 (optional): registration r uses src_count[r] points from src_offset[r] on (gapped layout, e.g. the
    output of voxel_down_sample, no repacking).  Returns (transform [R,4,4] fp64, info [R,4] fp64 = fitness, rmse, iterations, n_corr)."""
    require_cuda(source, target, src_offset, tgt_offset)
    source = _c(source, torch.float64); target = _c(target, torch.float64)
    src_offset = _c(src_offset, torch.int32); tgt_offset = _c(tgt_offset, torch.int32)
    R = src_offset.numel() - 1
    S, T = source.shape[0], target.shape[0]
    dev = source.device
    lib = _lib.load()
    work = torch.empty((int(lib.ape_icp_work_bytes(S, T)) + 7) // 8, dtype=torch.float64, device=dev)
    tf = torch.empty((R, 4, 4), dtype=torch.float64, device=dev)
    info = torch.empty((R, 4), dtype=torch.float64, device=dev)
    if init is not None:
        init = _c(init, torch.float64)
    if src_count is not None:
        src_count = _c(src_count, torch.int32)
    check(lib.ape_icp_p2p_ex(ptr(source), ptr(src_offset), ptr(src_count), ptr(target), ptr(tgt_offset), R, S, T, float(threshold),
                             float(rel_fitness), float(rel_rmse), int(max_iter), ptr(init), ptr(tf), ptr(info), ptr(work),
                             stream_ptr()), 'ape_icp_p2p_ex')
    return tf, info


def reconstruct_run(points, offset_host, voxel_size, threshold, rel_fitness=1e-2, rel_rmse=1e-2, max_iter=100):
    """The sequential register-and-merge loop of create_pointcloud.py:286-312 in ONE call without a host synchronisation.
    points [P,3] fp64 (the views' surfaces packed), offset_host [V+1] numpy int32 -> (cloud [P,3] fp64, count [1] int32,
    status [1] int32), the cloud occupies cloud[:count]; status != 0: an intermediate cloud was too large for the batched
    voxel kernel (run the loop view by view instead)."""
    import numpy as np
    require_cuda(points)
    points = _c(points, torch.float64)
    oh = np.ascontiguousarray(offset_host, dtype=np.int32)
    V = oh.size - 1
    if V < 0 or int(oh[-1]) != points.shape[0]:
        raise ValueError('reconstruct_run: offset_host must be [n_views + 1] and end at the number of points')
    dev = points.device
    lib = _lib.load()
    P = points.shape[0]
    work = torch.empty((int(lib.ape_reconstruct_work_bytes(P)) + 7) // 8, dtype=torch.float64, device=dev)
    out = torch.empty((P, 3), dtype=torch.float64, device=dev)
    cnt = torch.empty((1,), dtype=torch.int32, device=dev)
    status = torch.empty((1,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib.ape_reconstruct_run(ptr(points), oh.ctypes.data, V, float(voxel_size), float(threshold), float(rel_fitness),
                                      float(rel_rmse), int(max_iter), ptr(out), ptr(cnt), ptr(status), ptr(work), stream_ptr()),
              'ape_reconstruct_run')
    return out, cnt, status


VOXEL_MAX_POINTS = 16384


@_on_tensor_device
def voxel_down_sample(points, offset, voxel_size, offset_host=None, max_cloud_points=None):
    """a5/a6.  Ragged batch points [P,3] fp64, offset [C+1] int32 -> (out_points [P,3] fp64 (cloud c occupies
    out[offset[c] : offset[c]+counts[c]]), counts [C] int32).
    Clouds of up to 16 384 points go through ONE batched launch (shared-memory sort); larger ones through the
    global-memory path (ape_voxel_down_sample_large), one call per such cloud, same result bit for bit -- no cloud size
    fails.  Routing needs the sizes on the host: pass `offset_host` (numpy) when it is at hand, or `max_cloud_points`
    (an upper bound the caller guarantees) to skip the check; otherwise, and only when P > 16 384, `offset` is copied to
    the host once (a synchronisation)."""
    require_cuda(points, offset)
    points = _c(points, torch.float64); offset = _c(offset, torch.int32)
    C = offset.numel() - 1
    out = torch.empty_like(points)
    counts = torch.empty((C,), dtype=torch.int32, device=points.device)
    lib = _lib.load()
    check(lib.ape_voxel_down_sample(ptr(points), ptr(offset), C, float(voxel_size), ptr(out), ptr(counts), stream_ptr()),
          'ape_voxel_down_sample')
    if points.shape[0] > VOXEL_MAX_POINTS and not (max_cloud_points is not None and max_cloud_points <= VOXEL_MAX_POINTS):
        oh = offset_host if offset_host is not None else offset.cpu().numpy()
        for c in range(C):
            n = int(oh[c + 1]) - int(oh[c])
            if n > VOXEL_MAX_POINTS:
                lo = int(oh[c])
                check(lib.ape_voxel_down_sample_large(points[lo:lo + n].data_ptr(), n, float(voxel_size), out[lo:lo + n].data_ptr(),
                                                      counts[c:c + 1].data_ptr(), stream_ptr()), 'ape_voxel_down_sample_large')
    return out, counts


def _max_cloud(offset_host):
    import numpy as np
    return int(np.diff(offset_host).max()) if len(offset_host) > 1 else 0


@_on_tensor_device
def radius_outlier(points, offset, nb_points, radius, offset_host=None):
    """f-1.  Ragged batch points [P,3] fp64 + offset [C+1] int32 -> keep [P] uint8 (pcd.remove_radius_outlier)."""
    require_cuda(points, offset)
    points = _c(points, torch.float64); offset = _c(offset, torch.int32)
    oh = offset_host if offset_host is not None else offset.cpu().numpy()
    keep = torch.zeros((points.shape[0],), dtype=torch.uint8, device=points.device)
    check(_lib.load().ape_radius_outlier(ptr(points), ptr(offset), offset.numel() - 1, _max_cloud(oh), int(nb_points), float(radius),
                                         ptr(keep), None, stream_ptr()), 'ape_radius_outlier')
    return keep


@_on_tensor_device
def mahalanobis(points, offset, want_dist=True):
    """f-1.  -> (dist [P] fp64 or None, std [C] fp64 = np.std(|dist|) per cloud) (pcd.compute_mahalanobis_distance)."""
    require_cuda(points, offset)
    points = _c(points, torch.float64); offset = _c(offset, torch.int32)
    C = offset.numel() - 1
    dist = torch.zeros((points.shape[0],), dtype=torch.float64, device=points.device) if want_dist else None
    std = torch.zeros((C,), dtype=torch.float64, device=points.device)
    check(_lib.load().ape_mahalanobis(ptr(points), ptr(offset), C, ptr(dist), ptr(std), stream_ptr()), 'ape_mahalanobis')
    return dist, std


@_on_tensor_device
def statistical_outlier(points, offset, nb_neighbors, std_ratio, offset_host=None):
    """f-1.  std_ratio: python float or a [C] fp64 device tensor (one ratio per cloud) ->
    (keep [P] uint8, avg_dist [P] fp64, threshold [C] fp64) (pcd.remove_statistical_outlier)."""
    require_cuda(points, offset)
    points = _c(points, torch.float64); offset = _c(offset, torch.int32)
    oh = offset_host if offset_host is not None else offset.cpu().numpy()
    C = offset.numel() - 1
    keep = torch.zeros((points.shape[0],), dtype=torch.uint8, device=points.device)
    avg = torch.zeros((points.shape[0],), dtype=torch.float64, device=points.device)
    thr = torch.zeros((C,), dtype=torch.float64, device=points.device)
    dev_ratio = _c(std_ratio, torch.float64) if isinstance(std_ratio, torch.Tensor) else None
    check(_lib.load().ape_statistical_outlier(ptr(points), ptr(offset), C, _max_cloud(oh), int(nb_neighbors), ptr(dev_ratio),
                                              0.0 if dev_ratio is not None else float(std_ratio), ptr(keep), ptr(avg), ptr(thr),
                                              stream_ptr()), 'ape_statistical_outlier')
    return keep, avg, thr


@_on_tensor_device
def compact_points(points, offset, keep, want_index=False):
    """Ordered per-cloud compaction -> (out_points [P,3] (cloud c at offset[c] : offset[c]+counts[c]), counts [C][, index [P]])."""
    require_cuda(points, offset, keep)
    points = _c(points, torch.float64); offset = _c(offset, torch.int32); keep = _c(keep, torch.uint8)
    C = offset.numel() - 1
    out = torch.empty_like(points)
    counts = torch.zeros((C,), dtype=torch.int32, device=points.device)
    index = torch.empty((points.shape[0],), dtype=torch.int32, device=points.device) if want_index else None
    check(_lib.load().ape_compact_points(ptr(points), ptr(offset), ptr(keep), C, ptr(out), ptr(counts), ptr(index), stream_ptr()),
          'ape_compact_points')
    return (out, counts, index) if want_index else (out, counts)


@_on_tensor_device
def pose_select(pred_r, pred_t, pred_c, cloud, want_new_points=True):
    """a8/a9.  pred_r [B,N,4], pred_t [B,N,3], pred_c [B,N(,1)], cloud [B,N,3] fp32 ->
    dict(which_max [B] int32, my_r [B,4], my_t [B,3], new_points [B,N,3] or None, pose [B,7] fp64)."""
    require_cuda(pred_r, pred_t, pred_c, cloud)
    pred_r = _c(pred_r, torch.float32); pred_t = _c(pred_t, torch.float32)
    pred_c = _c(pred_c, torch.float32); cloud = _c(cloud, torch.float32)
    B, N = pred_r.shape[0], pred_r.shape[1]
    dev = pred_r.device
    wm = torch.empty((B,), dtype=torch.int32, device=dev)
    my_r = torch.empty((B, 4), dtype=torch.float32, device=dev)
    my_t = torch.empty((B, 3), dtype=torch.float32, device=dev)
    newp = torch.empty((B, N, 3), dtype=torch.float32, device=dev) if want_new_points else None
    pose = torch.empty((B, 7), dtype=torch.float64, device=dev)
    check(_lib.load().ape_pose_select(ptr(pred_r), ptr(pred_t), ptr(pred_c), ptr(cloud), B, N, ptr(wm), ptr(my_r), ptr(my_t),
                                      ptr(newp), ptr(pose), stream_ptr()), 'ape_pose_select')
    return dict(which_max=wm, my_r=my_r, my_t=my_t, new_points=newp, pose=pose)


@_on_tensor_device
def pose_compose(pose_in, r2, t2, cloud=None):
    """a11.  pose_in [B,7] fp64, r2 [B,4] fp32, t2 [B,3] fp32 -> pose_out [B,7] fp64
    (and next_points [B,N,3] if cloud [B,N,3] is given)."""
    require_cuda(pose_in, r2, t2, cloud)
    pose_in = _c(pose_in, torch.float64); r2 = _c(r2, torch.float32); t2 = _c(t2, torch.float32)
    B = pose_in.shape[0]
    out = torch.empty_like(pose_in)
    nxt, N = None, 0
    if cloud is not None:
        cloud = _c(cloud, torch.float32)
        N = cloud.shape[1]
        nxt = torch.empty_like(cloud)
    check(_lib.load().ape_pose_compose(ptr(pose_in), ptr(r2), ptr(t2), B, ptr(out), ptr(cloud), N, ptr(nxt), stream_ptr()),
          'ape_pose_compose')
    return (out, nxt) if cloud is not None else out


# ------------------------------------------------------------------------------------------ networks
NET_POSENET, NET_REFINER = 0, 1
GEMM_TCGEN05, GEMM_SIMT, GEMM_TCGEN05_V1, GEMM_TCGEN05_PAIR, GEMM_TCGEN05_B2B = 0, 1, 2, 3, 4

_POSENET_ORDER = ['feat.conv1', 'feat.e_conv1', 'feat.conv2', 'feat.e_conv2', 'feat.conv5', 'feat.conv6',
                  'conv1_r', 'conv1_t', 'conv1_c', 'conv2_r', 'conv2_t', 'conv2_c', 'conv3_r', 'conv3_t', 'conv3_c',
                  'conv4_r', 'conv4_t', 'conv4_c']
_REFINER_ORDER = ['feat.conv1', 'feat.e_conv1', 'feat.conv2', 'feat.e_conv2', 'feat.conv5', 'feat.conv6',
                  'conv1_r', 'conv1_t', 'conv2_r', 'conv2_t', 'conv3_r', 'conv3_t']


class NetHandle:
    """Owns an `ape_net` (split-bf16 weights + activation workspace on the current CUDA device).

    state_dict: reference-shaped tensors/arrays (DenseFusion/lib/network.py:74-91 or :139-183);
    extra keys (e.g. the colour encoder `cnn.*`) are ignored."""

    def __init__(self, kind, state_dict, num_obj, max_batch, max_points, device=None):
        import ctypes
        import numpy as np
        if not torch.cuda.is_available():
            raise _lib.ApeError('no CUDA device: the B200 path has no CPU fallback')
        self.device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        if self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        lib = _lib.load()
        order = _POSENET_ORDER if kind == NET_POSENET else _REFINER_ORDER
        host = []
        for name in order:
            for suffix in ('.weight', '.bias'):
                v = state_dict[name + suffix]
                if isinstance(v, torch.Tensor):
                    v = v.detach().to('cpu', torch.float32).numpy()
                host.append(np.ascontiguousarray(v, dtype=np.float32).reshape(-1))
        arr = (ctypes.c_void_p * len(host))(*[h.ctypes.data for h in host])
        out = ctypes.c_void_p()
        with torch.cuda.device(self.device):                  # weights and workspace live on the handle's device
            check(lib.ape_net_create(kind, arr, len(host), int(num_obj), int(max_batch), int(max_points), ctypes.byref(out)),
                  'ape_net_create')
        self._h = out
        self.kind, self.num_obj, self.max_batch, self.max_points = kind, num_obj, max_batch, max_points

    def set_gemm(self, impl):
        check(_lib.load().ape_net_set_gemm(self._h, impl), 'ape_net_set_gemm')

    def set_passes(self, masks):
        """Split-bf16 products per tensor-core layer (conv2, conv5, conv6, heads1, heads2, heads3): 7 = all three,
        6 = activation low half dropped, 5 = weight low half dropped, 4 = plain bf16."""
        import ctypes
        m = list(masks) + [7] * (6 - len(masks))
        arr = (ctypes.c_int * 6)(*m)
        check(_lib.load().ape_net_set_passes(self._h, arr), 'ape_net_set_passes')

    def get_passes(self):
        import ctypes
        arr = (ctypes.c_int * 6)()
        check(_lib.load().ape_net_get_passes(self._h, arr), 'ape_net_get_passes')
        return list(arr)

    def close(self):
        if getattr(self, '_h', None):
            _lib.load().ape_net_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @_on_tensor_device
    def posenet_forward(self, out_img, cloud, choose, obj, gathered=False):
        """out_img [B,32,hw] / [B,32,H,W] (contiguous, or 4-D in torch.channels_last memory format), cloud [B,N,3],
        choose [B,N] (or [B,1,N]) int64, obj [B] (or [B,1]) int64 -> pred_r [B,N,4], pred_t [B,N,3], pred_c [B,N,1], emb [B,32,N].
        gathered=True: out_img IS emb [B,32,N] (already gathered, choose ignored) and is returned as emb."""
        require_cuda(out_img, cloud, choose, obj)
        B, N = cloud.shape[0], cloud.shape[1]
        img, hw, layout = _emb_input(out_img, B, N, gathered)
        cloud = _c(cloud, torch.float32); obj = _c(obj, torch.int64).reshape(B)
        choose = None if gathered else _c(choose, torch.int64).reshape(B, N)
        dev = cloud.device
        r = torch.empty((B, N, 4), dtype=torch.float32, device=dev); t = torch.empty((B, N, 3), dtype=torch.float32, device=dev)
        c = torch.empty((B, N, 1), dtype=torch.float32, device=dev)
        emb = img if gathered else torch.empty((B, 32, N), dtype=torch.float32, device=dev)
        check(_lib.load().ape_posenet_forward_ex(self._h, img.data_ptr(), hw, layout, ptr(cloud), ptr(choose), ptr(obj), B, N,
                                                 ptr(r), ptr(t), ptr(c), None if gathered else ptr(emb), stream_ptr()),
              'ape_posenet_forward_ex')
        return r, t, c, emb

    @_on_tensor_device
    def refiner_forward(self, new_points, emb, obj):
        """new_points [B,N,3], emb [B,32,N], obj [B] -> r2 [B,4], t2 [B,3]"""
        require_cuda(new_points, emb, obj)
        B, N = new_points.shape[0], new_points.shape[1]
        new_points = _c(new_points, torch.float32); emb = _c(emb, torch.float32); obj = _c(obj, torch.int64).reshape(B)
        r2 = torch.empty((B, 4), dtype=torch.float32, device=new_points.device)
        t2 = torch.empty((B, 3), dtype=torch.float32, device=new_points.device)
        check(_lib.load().ape_refiner_forward(self._h, ptr(new_points), ptr(emb), ptr(obj), B, N, ptr(r2), ptr(t2), stream_ptr()),
              'ape_refiner_forward')
        return r2, t2


EMB_NCHW, EMB_NHWC, EMB_GATHERED = 0, 1, 2


def _emb_input(out_img, B, N, gathered):
    """-> (fp32 tensor whose data_ptr is handed over, hw, APE_EMB_* layout) for an encoder map / a gathered embedding."""
    if out_img.dtype != torch.float32:
        out_img = out_img.float()
    if gathered:
        assert tuple(out_img.shape) == (B, 32, N), 'gathered embedding must be [B,32,N]'
        return out_img.contiguous(), 0, EMB_GATHERED
    if out_img.dim() == 4 and not out_img.is_contiguous() and out_img.is_contiguous(memory_format=torch.channels_last):
        return out_img, out_img.shape[2] * out_img.shape[3], EMB_NHWC
    out_img = out_img.contiguous().reshape(B, 32, -1)
    return out_img, out_img.shape[2], EMB_NCHW


@_on_tensor_device
def gather_emb(out_img, choose, out=None):
    """network.py:100-102 as a stand-alone kernel.  out_img [B,32,hw] / [B,32,H,W] fp32: a CUDA tensor, or a PINNED host
    tensor (read zero-copy over PCIe: only the sampled columns cross the bus); contiguous or channels_last.
    choose [B,N] int64 CUDA -> emb [B,32,N] fp32 CUDA."""
    require_cuda(choose)
    if not out_img.is_cuda and not out_img.is_pinned():
        raise _lib.ApeError('gather_emb: a host encoder map must be pinned (torch.Tensor.pin_memory) for zero-copy access')
    B = out_img.shape[0]
    choose = _c(choose, torch.int64).reshape(B, -1)
    N = choose.shape[1]
    img, hw, layout = _emb_input(out_img, B, N, False)
    emb = out if out is not None else torch.empty((B, 32, N), dtype=torch.float32, device=choose.device)
    check(_lib.load().ape_gather_emb(img.data_ptr(), hw, layout, ptr(choose), B, N, ptr(emb), stream_ptr()), 'ape_gather_emb')
    return emb


def host_gather_begin(out_img, choose, emb_host, obj_begin=0, obj_end=None, threads=0):
    """Start gathering objects [obj_begin, obj_end) of a HOST encoder map into the (pinned) staging tensor emb_host
    [B,32,N] on the library's thread pool; returns immediately, host_gather_wait() blocks until it is complete."""
    assert not out_img.is_cuda and not choose.is_cuda and not emb_host.is_cuda
    B = out_img.shape[0]
    choose = choose.reshape(B, -1)
    assert choose.dtype == torch.int64 and choose.is_contiguous() and emb_host.is_contiguous() and emb_host.dtype == torch.float32
    N = choose.shape[1]
    img, hw, layout = _emb_input(out_img, B, N, False)
    check(_lib.load().ape_host_gather_begin(img.data_ptr(), hw, layout, choose.data_ptr(), int(obj_begin),
                                            int(B if obj_end is None else obj_end), N, emb_host.data_ptr(), int(threads)),
          'ape_host_gather_begin')
    return img                                                   # keep alive until host_gather_wait()


def host_gather_wait():
    check(_lib.load().ape_host_gather_wait(), 'ape_host_gather_wait')


@_on_tensor_device
def pose_pipeline(est, ref, out_img, cloud, choose, obj, iterations=2, canonical=True, out=None, gathered=False):
    """Whole option-6 geometry block: -> (poses [B,7] fp64 (wxyz, t), which_max [B] int32).  No host sync.
    out_img: encoder map [B,32,hw] / [B,32,H,W] (contiguous or channels_last), or with gathered=True the already
    gathered emb [B,32,N] (choose may then be None)."""
    require_cuda(out_img, cloud, choose, obj)
    B, N = cloud.shape[0], cloud.shape[1]
    img, hw, layout = _emb_input(out_img, B, N, gathered)
    cloud = _c(cloud, torch.float32); obj = _c(obj, torch.int64).reshape(B)
    choose = None if gathered else _c(choose, torch.int64).reshape(B, N)
    poses = out if out is not None else torch.empty((B, 7), dtype=torch.float64, device=cloud.device)
    wm = torch.empty((B,), dtype=torch.int32, device=cloud.device)
    check(_lib.load().ape_pose_pipeline_ex(est._h, ref._h if ref is not None else None, img.data_ptr(), hw, layout, ptr(cloud),
                                           ptr(choose), ptr(obj), B, N, int(iterations), int(bool(canonical)), ptr(poses), ptr(wm),
                                           stream_ptr()), 'ape_pose_pipeline_ex')
    return poses, wm


# ------------------------------------------------------------------------------------------ refiner training (a16)
@_on_tensor_device
def refine_loss(pred_r, pred_t, model_points, target, points=None, symmetric=None, want_grad=True, want_next=True):
    """Loss_refine forward + backward (lib/loss_refiner.py:12-64) for B objects.  pred_r [B,4], pred_t [B,3],
    model_points / target [B,M,3], points [B,N,3], symmetric [B] bool/uint8 ->
    dict(dis [B], d_r [B,4], d_t [B,3], new_points [B,N,3], new_target [B,M,3])."""
    require_cuda(pred_r, pred_t, model_points, target, points)
    pred_r = _c(pred_r, torch.float32); pred_t = _c(pred_t, torch.float32)
    model_points = _c(model_points, torch.float32); target = _c(target, torch.float32)
    B, M = model_points.shape[0], model_points.shape[1]
    dev = pred_r.device
    sym = _c(symmetric.to(torch.uint8), torch.uint8) if symmetric is not None else None
    dis = torch.empty((B,), dtype=torch.float32, device=dev)
    d_r = torch.empty((B, 4), dtype=torch.float32, device=dev) if want_grad else None
    d_t = torch.empty((B, 3), dtype=torch.float32, device=dev) if want_grad else None
    newp = newt = None
    N = 0
    if want_next:
        points = _c(points, torch.float32)
        N = points.shape[1]
        newp = torch.empty_like(points); newt = torch.empty_like(target)
    check(_lib.load().ape_refine_loss(ptr(pred_r), ptr(pred_t), ptr(model_points), ptr(target), M, ptr(points), N, ptr(sym), B,
                                      ptr(dis), ptr(d_r), ptr(d_t), ptr(newp), ptr(newt), stream_ptr()), 'ape_refine_loss')
    return dict(dis=dis, d_r=d_r, d_t=d_t, new_points=newp, new_target=newt)


@_on_tensor_device
def adam_step(params, grads, exp_avg, exp_avg_sq, step, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0):
    """torch.optim.Adam update on flat fp32 device vectors (train.py:149)."""
    require_cuda(params, grads, exp_avg, exp_avg_sq)
    check(_lib.load().ape_adam_step(ptr(params), ptr(grads), ptr(exp_avg), ptr(exp_avg_sq), params.numel(), float(lr),
                                    float(betas[0]), float(betas[1]), float(eps), int(step), float(grad_scale), stream_ptr()),
          'ape_adam_step')


def refiner_trainer_layout(num_obj):
    """-> (flat parameter count, {reference state_dict key: (offset, shape)}) of the trainer's flat vector."""
    import ctypes
    off = (ctypes.c_int64 * 24)()
    total = _lib.load().ape_refiner_trainer_layout(int(num_obj), off)
    if total <= 0:
        raise _lib.ApeError('ape_refiner_trainer_layout failed')
    shapes = {'feat.conv1': (64, 3, 1), 'feat.e_conv1': (64, 32, 1), 'feat.conv2': (128, 64, 1), 'feat.e_conv2': (128, 64, 1),
              'feat.conv5': (512, 384, 1), 'feat.conv6': (1024, 512, 1), 'conv1_r': (512, 1024), 'conv1_t': (512, 1024),
              'conv2_r': (128, 512), 'conv2_t': (128, 512), 'conv3_r': (num_obj * 4, 128), 'conv3_t': (num_obj * 3, 128)}
    table = {}
    for i, name in enumerate(_REFINER_ORDER):
        table[name + '.weight'] = (int(off[2 * i]), shapes[name])
        table[name + '.bias'] = (int(off[2 * i + 1]), (shapes[name][0],))
    return int(total), table


class RefinerTrainerHandle:
    """Owns an `ape_trainer`: flat fp32 parameter / gradient vectors (torch tensors, so NCCL can all-reduce the
    gradient in place) plus the bf16 weight copies and the activation workspace of the training forward."""

    def __init__(self, state_dict, num_obj, max_batch, max_points, device=None):
        import ctypes
        if not torch.cuda.is_available():
            raise _lib.ApeError('no CUDA device: the B200 path has no CPU fallback')
        self.device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        total, self.table = refiner_trainer_layout(num_obj)
        self.params = torch.zeros((total,), dtype=torch.float32, device=self.device)
        self.grads = torch.zeros_like(self.params)
        self.num_obj, self.max_batch, self.max_points = num_obj, max_batch, max_points
        self._h = None
        self.load_state_dict(state_dict, sync=False)
        out = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(_lib.load().ape_refiner_trainer_create(ptr(self.params), ptr(self.grads), int(num_obj), int(max_batch),
                                                         int(max_points), ctypes.byref(out)), 'ape_refiner_trainer_create')
        self._h = out

    def view(self, key, flat=None):
        off, shape = self.table[key]
        n = 1
        for s in shape:
            n *= s
        return (self.params if flat is None else flat)[off:off + n].view(shape)

    def load_state_dict(self, state_dict, sync=True):
        for key in self.table:
            v = state_dict[key]
            if not isinstance(v, torch.Tensor):
                v = torch.as_tensor(v)
            self.view(key).copy_(v.detach().to(self.device, torch.float32).reshape(self.table[key][1]))
        if sync and self._h is not None:
            self.sync_weights()

    def state_dict(self):
        return {key: self.view(key).detach().clone() for key in self.table}

    def sync_weights(self):
        check(_lib.load().ape_refiner_trainer_sync_weights(self._h, stream_ptr()), 'ape_refiner_trainer_sync_weights')

    def forward(self, new_points, emb, obj):
        require_cuda(new_points, emb, obj)
        B, N = new_points.shape[0], new_points.shape[1]
        new_points = _c(new_points, torch.float32); emb = _c(emb, torch.float32); obj = _c(obj, torch.int64).reshape(B)
        r2 = torch.empty((B, 4), dtype=torch.float32, device=new_points.device)
        t2 = torch.empty((B, 3), dtype=torch.float32, device=new_points.device)
        check(_lib.load().ape_refiner_trainer_forward(self._h, ptr(new_points), ptr(emb), ptr(obj), B, N, ptr(r2), ptr(t2),
                                                      stream_ptr()), 'ape_refiner_trainer_forward')
        return r2, t2

    def backward(self, new_points, emb, obj, d_r, d_t):
        """Accumulates into self.grads; must follow the forward() of the same inputs."""
        B, N = new_points.shape[0], new_points.shape[1]
        new_points = _c(new_points, torch.float32); emb = _c(emb, torch.float32); obj = _c(obj, torch.int64).reshape(B)
        d_r = _c(d_r, torch.float32); d_t = _c(d_t, torch.float32)
        check(_lib.load().ape_refiner_trainer_backward(self._h, ptr(new_points), ptr(emb), ptr(obj), B, N, ptr(d_r), ptr(d_t),
                                                       stream_ptr()), 'ape_refiner_trainer_backward')

    def step(self, points, emb, obj, target, model_points, symmetric=None, iterations=2, zero_grad=True, out=None):
        """Accumulation phase of one optimizer step in one C call (no Python between the launches): -> dis [iterations, B]."""
        require_cuda(points, emb, obj, target, model_points)
        B, N, M = points.shape[0], points.shape[1], model_points.shape[1]
        points = _c(points, torch.float32); emb = _c(emb, torch.float32); obj = _c(obj, torch.int64).reshape(B)
        target = _c(target, torch.float32); model_points = _c(model_points, torch.float32)
        sym = _c(symmetric.to(torch.uint8), torch.uint8) if symmetric is not None else None
        dis = out if out is not None else torch.empty((iterations, B), dtype=torch.float32, device=points.device)
        check(_lib.load().ape_refiner_trainer_step(self._h, ptr(points), ptr(emb), ptr(obj), ptr(target), ptr(model_points), ptr(sym),
                                                   B, N, M, int(iterations), int(bool(zero_grad)), ptr(dis), stream_ptr()),
              'ape_refiner_trainer_step')
        return dis

    def wait_bulk(self, side_stream):
        """Make `side_stream` (torch.cuda.Stream) wait until the tail block grads[bulk_begin:] of the step just enqueued is
        final (conv6 + heads, 89 % of the vector); returns bulk_begin."""
        import ctypes
        lo = ctypes.c_int64(0)
        check(_lib.load().ape_refiner_trainer_wait_bulk(self._h, side_stream.cuda_stream, ctypes.byref(lo)), 'ape_refiner_trainer_wait_bulk')
        return int(lo.value)

    def adam(self, exp_avg, exp_avg_sq, step, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0):
        """Adam on the flat vectors + refresh of the bf16 weight copies (one C call)."""
        check(_lib.load().ape_refiner_trainer_adam(self._h, ptr(exp_avg), ptr(exp_avg_sq), float(lr), float(betas[0]), float(betas[1]),
                                                   float(eps), int(step), float(grad_scale), stream_ptr()), 'ape_refiner_trainer_adam')

    def close(self):
        if getattr(self, '_h', None):
            _lib.load().ape_refiner_trainer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
