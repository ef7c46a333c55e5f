"""Hot loop of object-cloud reconstruction (reference: pc_reconstruction/create_pointcloud.py:232-317, called from
`load_point_cloud` :181 for main.py option 4).

Per selected view the reference does: read meta / depth / label -> `get_surface` (per-pixel Python loop + open3d filters)
-> `icp_regression` against the cloud built so far -> transform, concatenate, voxel grid.  Here:

  get_surfaces_batch   every view of a run in ONE pass of batched launches (back-projection, voxel grid, radius /
                       statistical outlier filters over a ragged batch) -- the views are independent up to this point;
  reconstruct_run      the sequential part (each view registers to the cloud accumulated so far: "replicas only",
                       SURVEY 8e) on the persistent ICP kernel.

View selection, plots and the .pcd/.ply writers stay outside the graft (component 5); `formats.py` covers the
files on either side.  Millimetres, fp64, no CPU fallback."""
import numpy as np
import torch

from .. import ops
from .open3d_utils import PointCloud, icp_regression


def _repack(flat, starts, counts):
    """Gapped ragged layout (cloud c at flat[starts[c] : starts[c] + counts[c]]) -> contiguous (flat', offsets int32 host)."""
    counts = np.asarray(counts, np.int64)
    off = np.zeros(len(counts) + 1, np.int32)
    off[1:] = np.cumsum(counts)
    if len(counts) == 0 or off[-1] == 0:
        return flat[:0].contiguous(), off
    idx = np.concatenate([np.arange(s, s + c) for s, c in zip(np.asarray(starts, np.int64), counts)])
    return flat[torch.from_numpy(idx).to(flat.device)].contiguous(), off


def get_surfaces_batch(label, depth, cam, robot2cam, min_friends, min_dist, nb_neighbors, voxel_size, capacity=None,
                       outlier_filters=True):
    """`get_surface` (open3d_utils.py:171-213) for V views at once.  label [V,H,W] uint8, depth [V,H,W] uint16/int16
    storage, cam [V,4] fp64 (ppx,ppy,fx,fy), robot2cam [V,4,4] fp64 -- CUDA tensors (see formats.FrameBatchLoader).
    Returns a list of V PointClouds (possibly empty ones, which the caller skips as :286-287 does)."""
    V = label.shape[0]
    dev = label.device
    cap = int(capacity or int((label != 0).reshape(V, -1).sum(dim=1).max().item()) or 1)
    pts, _, cnt = ops.surface_backproject(label, depth, cam, robot2cam, capacity=cap, want_pixels=False)
    cnt_h = np.minimum(cnt.cpu().numpy(), cap)
    flat, off = _repack(pts.reshape(-1, 3), np.arange(V) * cap, cnt_h)
    if voxel_size and off[-1] > 0:
        out, vc = ops.voxel_down_sample(flat, torch.from_numpy(off).to(dev), float(voxel_size))
        vc_h = vc.cpu().numpy()
        if (vc_h < 0).any():
            raise ops._lib.ApeError('voxel_down_sample failed for a view (status %d)' % vc_h.min())
        flat, off = _repack(out, off[:-1], vc_h)
    if outlier_filters and off[-1] > 0:
        off_d = torch.from_numpy(off).to(dev)
        keep = ops.radius_outlier(flat, off_d, min_friends, min_dist, off)
        out, kc = ops.compact_points(flat, off_d, keep)
        flat, off = _repack(out, off[:-1], kc.cpu().numpy())
        if off[-1] > 0:
            off_d = torch.from_numpy(off).to(dev)
            _, ratio = ops.mahalanobis(flat, off_d, want_dist=False)           # per-view std of |Mahalanobis|, stays on device
            keep, _, _ = ops.statistical_outlier(flat, off_d, nb_neighbors, ratio, off)
            # a view with fewer than two points has no defined std: the reference would produce NaN and drop everything
            out, kc = ops.compact_points(flat, off_d, keep)
            flat, off = _repack(out, off[:-1], kc.cpu().numpy())
    return [PointCloud(flat[off[v]:off[v + 1]].clone()) for v in range(V)]


def reconstruct_run(surfaces, voxel_size, threshold, global_regression=False, icp_point2point=True, icp_point2plane=False):
    """create_pointcloud.py:286-312: first non-empty surface starts the cloud; every further one is registered to it
    (`icp_regression`), transformed, concatenated IN FRONT of the target and voxel-down-sampled."""
    cloud = None
    for source in surfaces:
        if len(source) == 0:
            continue
        if cloud is None:
            cloud = PointCloud(source.points.clone())
            continue
        target, source, T = icp_regression(cloud, source, voxel_size=voxel_size, threshold=threshold,
                                           global_regression=global_regression, icp_point2point=icp_point2point,
                                           icp_point2plane=icp_point2plane)
        source = source.transform(T)
        cloud = PointCloud(torch.cat((source.points, target.points))).voxel_down_sample(voxel_size)
    return cloud if cloud is not None else PointCloud()


def rotate_about_center(cloud, R):
    """`point_cloud.rotate(R=point_cloud_tf, center=True)` (:320, open3d 0.9: rotation about the cloud's mean)."""
    R = torch.as_tensor(np.asarray(R, np.float64), device=cloud.points.device)
    c = cloud.points.mean(dim=0, keepdim=True)
    cloud.points = ((cloud.points - c) @ R.T + c).contiguous()
    return cloud
