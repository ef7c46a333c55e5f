"""Drop-in for the hot functions of pc_reconstruction/open3d_utils.py (option 4, "Create Pose labels").

  get_surface     :171-213  masked depth -> robot-frame cloud (+ voxel grid)       -> csrc/backproject.cu, voxel.cu
  icp_regression  :63-122   voxel grid on both clouds + point-to-point ICP          -> csrc/voxel.cu, icp.cu
  icp_regression_batch      the same for many (target, source) pairs in one launch (BASELINE config 4)

open3d is not a dependency: clouds are carried by `PointCloud`, a minimal stand-in for
o3d.geometry.PointCloud (points, transform, voxel_down_sample, get_center, translate) backed by an [n,3] fp64
CUDA tensor.  Units are millimetres as in the reference.  There is no CPU fallback.

  get_surface's filter chain :198-211 and align_point_clouds :125-168 (voxel grid -> remove_radius_outlier ->
  std of the Mahalanobis distances -> remove_statistical_outlier) run on csrc/filters.cu.
"""
import numpy as np
import torch

from .. import ops


def _dev():
    if not torch.cuda.is_available():
        raise ops._lib.ApeError('no CUDA device: the B200 path has no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


class PointCloud:
    """Minimal o3d.geometry.PointCloud stand-in: `.points` is an [n,3] fp64 CUDA tensor."""

    def __init__(self, points=None):
        if points is None:
            points = torch.zeros((0, 3), dtype=torch.float64, device=_dev())
        elif not isinstance(points, torch.Tensor):
            points = torch.from_numpy(np.ascontiguousarray(points, dtype=np.float64)).to(_dev())
        self.points = points.to(torch.float64).reshape(-1, 3).contiguous()

    def __len__(self):
        return self.points.shape[0]

    def numpy(self):
        return self.points.cpu().numpy()

    def get_center(self):
        return self.points.mean(dim=0).cpu().numpy()

    def translate(self, translation):
        self.points = self.points + torch.as_tensor(np.asarray(translation, np.float64), device=self.points.device)
        return self

    def transform(self, T):
        T = torch.as_tensor(np.asarray(T, np.float64), device=self.points.device)
        self.points = (self.points @ T[:3, :3].T + T[:3, 3]).contiguous()
        return self

    def voxel_down_sample(self, voxel_size):
        n = len(self)
        if n == 0:
            return PointCloud(self.points.clone())
        off = torch.tensor([0, n], dtype=torch.int32, device=self.points.device)
        out, cnt = ops.voxel_down_sample(self.points, off, voxel_size, offset_host=np.array([0, n]))
        c = int(cnt.cpu()[0])
        if c < 0:
            raise ops._lib.ApeError('voxel_down_sample: more voxels along an axis than the kernel indexes (%d points, status %d)' % (n, c))
        return PointCloud(out[:c].clone())

    def _offset(self):
        return torch.tensor([0, len(self)], dtype=torch.int32, device=self.points.device), np.array([0, len(self)], np.int32)

    def select_down_sample(self, keep):
        """Points with keep != 0, original order.  Returns (PointCloud, kept indices int64 tensor)."""
        idx = torch.nonzero(keep.to(torch.bool)).reshape(-1)
        return PointCloud(self.points[idx].clone()), idx

    def remove_radius_outlier(self, nb_points, radius):
        """o3d PointCloud.remove_radius_outlier -> (filtered cloud, kept indices)."""
        if len(self) == 0:
            return PointCloud(self.points.clone()), torch.zeros((0,), dtype=torch.int64, device=self.points.device)
        off, oh = self._offset()
        return self.select_down_sample(ops.radius_outlier(self.points, off, nb_points, radius, oh))

    def compute_mahalanobis_distance(self):
        """o3d PointCloud.compute_mahalanobis_distance -> [n] fp64 numpy (the reference takes np.std of it on the host)."""
        off, _ = self._offset()
        return ops.mahalanobis(self.points, off)[0].cpu().numpy()

    def remove_statistical_outlier(self, nb_neighbors, std_ratio):
        """o3d PointCloud.remove_statistical_outlier -> (filtered cloud, kept indices)."""
        if len(self) == 0:
            return PointCloud(self.points.clone()), torch.zeros((0,), dtype=torch.int64, device=self.points.device)
        off, oh = self._offset()
        keep, _, _ = ops.statistical_outlier(self.points, off, nb_neighbors, std_ratio, oh)
        return self.select_down_sample(keep)


def _filter_chain(surface, min_friends, min_dist, nb_neighbors):
    """open3d_utils.py:203-211 / :161-166: radius outliers, then statistical outliers whose ratio is the std of the
    Mahalanobis distances of the radius-filtered cloud -- all on the device, no host round trip in between."""
    surface, _ = surface.remove_radius_outlier(nb_points=min_friends, radius=min_dist)
    if len(surface) < 2:
        return surface
    off, oh = surface._offset()
    _, ratio = ops.mahalanobis(surface.points, off, want_dist=False)              # std_ratio = np.std(|dist|), stays on device
    keep, _, _ = ops.statistical_outlier(surface.points, off, nb_neighbors, ratio, oh)
    return surface.select_down_sample(keep)[0]


def get_surface(label, depth_frame, intr, robot2Cam_ft, min_friends=None, min_dist=None, nb_neighbors=None, voxel_size=None,
                outlier_filters=True):
    """open3d_utils.py:171-213.  label uint8 [H,W] (non-zero = object), depth_frame [H,W] raw depth (mm; integral
    values, as read from the 16-bit PNG), intr dict(ppx,ppy,fx,fy), robot2Cam_ft 4x4.  Returns a PointCloud.
    outlier_filters=False stops after the voxel grid (the back-projection + voxel part alone)."""
    if outlier_filters and (min_friends is None or min_dist is None or nb_neighbors is None or not voxel_size):
        raise ValueError('get_surface: min_friends, min_dist, nb_neighbors and voxel_size are required (open3d_utils.py:171)')
    dev = _dev()
    depth = np.asarray(depth_frame)
    d16 = depth.astype(np.uint16)
    if not np.array_equal(d16, depth):
        raise ValueError('get_surface: depth_frame must hold integral 16-bit sensor values')
    lab = torch.from_numpy(np.ascontiguousarray(label, dtype=np.uint8)[None]).to(dev)
    dep = torch.from_numpy(d16.view(np.int16)[None].copy()).to(dev)
    cam = torch.tensor([[intr['ppx'], intr['ppy'], intr['fx'], intr['fy']]], dtype=torch.float64, device=dev)
    r2c = torch.from_numpy(np.asarray(robot2Cam_ft, np.float64)[None].copy()).to(dev)
    cap = int(lab.shape[1] * lab.shape[2])
    n_hint = int((lab != 0).sum())                         # capacity = number of labelled pixels (upper bound)
    pts, _, cnt = ops.surface_backproject(lab, dep, cam, r2c, capacity=max(16, min(cap, n_hint)), want_pixels=False)
    surface = PointCloud(pts[0, :int(cnt.cpu()[0])].clone())
    if voxel_size:
        surface = surface.voxel_down_sample(voxel_size)
    if outlier_filters:
        surface = _filter_chain(surface, min_friends, min_dist, nb_neighbors)
    return surface


def align_point_clouds(point_clouds, min_friends, min_dist, nb_neighbors, plot=False, global_regression=False, icp_point2point=True,
                       icp_point2plane=False, voxel_size=5, threshold=50):
    """open3d_utils.py:125-168: register every further rotation run onto the first, merge, voxel grid, outlier filters."""
    target = point_clouds[0]
    for source in point_clouds[1:]:
        diff = np.array(source.get_center()) - np.array(target.get_center())
        if diff[1] > -30:                                                         # :141-143
            source.translate([0, -30 - diff[1], 0])
        target, source, init_tf = icp_regression(target, source, voxel_size=voxel_size, threshold=threshold,
                                                 global_regression=global_regression, icp_point2point=icp_point2point,
                                                 icp_point2plane=icp_point2plane)
        source = source.transform(init_tf)
        target = PointCloud(torch.cat((source.points, target.points)))            # :156-157 (source first)
        target = target.voxel_down_sample(voxel_size)
        target = _filter_chain(target, min_friends, min_dist, nb_neighbors)
    return target


def icp_regression_batch(targets, sources, voxel_size=5, threshold=100, max_iteration=100, relative_fitness=1e-2,
                         relative_rmse=1e-2):
    """Many independent registrations in one ICP launch.  targets/sources: lists of PointCloud.
    Returns (targets_down, sources_down, transforms [R,4,4] numpy fp64, info [R,4] numpy)."""
    dev = _dev()
    def pack(clouds):
        off = np.zeros(len(clouds) + 1, np.int32)
        off[1:] = np.cumsum([len(c) for c in clouds])
        flat = torch.cat([c.points for c in clouds]) if clouds else torch.zeros((0, 3), dtype=torch.float64, device=dev)
        return flat, off
    def down(clouds):
        flat, off = pack(clouds)
        out, cnt = ops.voxel_down_sample(flat, torch.from_numpy(off).to(dev), voxel_size, offset_host=off)
        cnt = cnt.cpu().numpy()
        if (cnt < 0).any():
            raise ops._lib.ApeError('voxel_down_sample failed for a cloud (status %s)' % cnt.min())
        return [PointCloud(out[off[i]:off[i] + cnt[i]].clone()) for i in range(len(clouds))]
    td, sd = down(targets), down(sources)
    tf, to = pack(td); sf, so = pack(sd)
    T, info = ops.icp_p2p(sf, torch.from_numpy(so).to(dev), tf, torch.from_numpy(to).to(dev), threshold, relative_fitness,
                          relative_rmse, max_iteration)
    return td, sd, T.cpu().numpy(), info.cpu().numpy()


def icp_regression(target, source, voxel_size=5, threshold=100, global_regression=False, icp_point2point=True,
                   icp_point2plane=True, plot=False):
    """open3d_utils.py:63-122 with the reference's signature AND defaults.  Only the configuration the reference runs is
    grafted (point-to-point; main.py:177-179 and create_labels.py:229-231 pass global_regression=False,
    icp_point2point=True, icp_point2plane=False).  The signature default icp_point2plane=True (:66-67) would run a
    point-to-plane refinement after the point-to-point one (:106-117): that is never silently skipped -- every call with
    icp_point2plane=True (or global_regression=True) raises, so pass icp_point2plane=False as the reference's callers do.
    Returns (target_down, source_down, T 4x4 numpy fp64)."""
    if global_regression or icp_point2plane:
        raise NotImplementedError('icp_regression: global (FPFH/RANSAC) registration and point-to-plane ICP are disabled by the '
                                  'reference configuration (main.py:177-179) and are not grafted; call with '
                                  'global_regression=False, icp_point2plane=False')
    td, sd, T, _ = icp_regression_batch([target], [source], voxel_size, threshold)
    return td[0], sd[0], (T[0] if icp_point2point else np.identity(4))
