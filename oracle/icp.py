"""Oracle (test infrastructure): point-to-point ICP + voxel grid, open3d-0.9.0 semantics.

The reference calls a third-party engine that is NOT vendored under the
reference tree and is NOT installable here: **open3d 0.9.0.0** (pinned in prose
only, README.md:44).  Call sites restated:
  * icp_regression                  pc_reconstruction/open3d_utils.py:63-122
      voxel_down_sample both clouds (:72-73 via :19-21), then
      registration_icp(source, target, threshold, I, PointToPoint(),
                       ICPConvergenceCriteria(1e-2, 1e-2, 100))   (:76-104)
  * merge                           pc_reconstruction/create_pointcloud.py:307-312
Published algorithm restated (Open3D v0.9.0 Registration.cpp, Eigen umeyama):
  correspondences: for every source point the nearest target point, kept when
  d^2 < r^2 (strict); fitness = |C|/|S|; inlier_rmse = sqrt(sum d^2 / |C|);
  update = umeyama(src_C, tgt_C, with_scaling=False) in fp64;
  T <- update @ T; the *working* cloud is transformed incrementally by `update`;
  correspondences are evaluated once before iteration 0 and after every update;
  stop when |d fitness| < 1e-2 and |d rmse| < 1e-2, or after max_iteration.
PARITY UNPINNED at the open3d boundary (no open3d here, no reference vectors);
the Kabsch step is pinned by transformations.py doctests (:909-925).
"""
import numpy as np


def kabsch_umeyama(src, dst):
    """Eigen::umeyama(src, dst, with_scaling=false) for 3xN fp64 correspondences
    given as [n,3] arrays.  Returns 4x4 fp64."""
    src = np.asarray(src, np.float64); dst = np.asarray(dst, np.float64)
    n = src.shape[0]
    mu_s = src.sum(axis=0) / n
    mu_d = dst.sum(axis=0) / n
    sigma = (dst - mu_d).T @ (src - mu_s) / n
    U, _, Vt = np.linalg.svd(sigma)
    S = np.ones(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        S[2] = -1.0
    R = U @ np.diag(S) @ Vt
    T = np.identity(4)
    T[:3, :3] = R
    T[:3, 3] = mu_d - R @ mu_s
    return T


def nn_within(src, tgt, radius, chunk=512):
    """Brute-force fp64 nearest neighbour, d2 = (dx*dx + dy*dy) + dz*dz, lowest
    index on exact ties, kept when d2 < radius^2 (strict).
    Returns (idx int64 [Ns] with -1 for no hit, d2 fp64 [Ns])."""
    src = np.asarray(src, np.float64); tgt = np.asarray(tgt, np.float64)
    idx = np.empty(len(src), np.int64); d2 = np.empty(len(src), np.float64)
    for s in range(0, len(src), chunk):
        p = src[s:s + chunk]
        dx = p[:, None, 0] - tgt[None, :, 0]
        dy = p[:, None, 1] - tgt[None, :, 1]
        dz = p[:, None, 2] - tgt[None, :, 2]
        dd = (dx * dx + dy * dy) + dz * dz
        j = dd.argmin(axis=1)
        idx[s:s + chunk] = j
        d2[s:s + chunk] = dd[np.arange(len(p)), j]
    miss = ~(d2 < float(radius) * float(radius))
    idx[miss] = -1
    return idx, d2


def nn_within_kdtree(src, tgt_tree, radius):
    """Same contract through scipy's cKDTree (used for the timed CPU baseline)."""
    d, j = tgt_tree.query(src, k=1, distance_upper_bound=radius)
    d2 = d * d
    miss = ~(d2 < float(radius) * float(radius))
    j = j.astype(np.int64)
    j[miss] = -1
    return j, d2


def _evaluate(work, tgt, radius, nn):
    idx, d2 = nn(work, tgt, radius)
    hit = idx >= 0
    n = int(hit.sum())
    if n == 0:
        return idx, hit, 0.0, 0.0
    fitness = n / float(len(work))
    rmse = float(np.sqrt(d2[hit].sum() / n))
    return idx, hit, fitness, rmse


def registration_icp_p2p(source, target, threshold, init=None,
                         relative_fitness=1e-2, relative_rmse=1e-2, max_iteration=100,
                         use_kdtree=False, return_info=False):
    """registration_icp(..., TransformationEstimationPointToPoint(), criteria).transformation
    (open3d_utils.py:98-104).  source/target fp64 [n,3].  Returns 4x4 fp64 (and, with
    return_info, dict(fitness, inlier_rmse, iterations, n_corr))."""
    source = np.asarray(source, np.float64); target = np.asarray(target, np.float64)
    T = np.identity(4) if init is None else np.array(init, np.float64)
    work = source.copy()
    if not np.array_equal(T, np.identity(4)):
        work = work @ T[:3, :3].T + T[:3, 3]
    if use_kdtree:
        from scipy.spatial import cKDTree
        tree = cKDTree(target)
        nn = lambda s, _t, r: nn_within_kdtree(s, tree, r)
    else:
        nn = nn_within
    idx, hit, fit, rmse = _evaluate(work, target, threshold, nn)
    it_done = 0
    for _ in range(max_iteration):
        if hit.any():
            upd = kabsch_umeyama(work[hit], target[idx[hit]])
        else:
            upd = np.identity(4)
        T = upd @ T
        work = work @ upd[:3, :3].T + upd[:3, 3]
        pfit, prmse = fit, rmse
        idx, hit, fit, rmse = _evaluate(work, target, threshold, nn)
        it_done += 1
        if abs(pfit - fit) < relative_fitness and abs(prmse - rmse) < relative_rmse:
            break
    if return_info:
        return T, dict(fitness=fit, inlier_rmse=rmse, iterations=it_done, n_corr=int(hit.sum()))
    return T


def voxel_down_sample(points, voxel_size):
    """open3d 0.9 PointCloud::VoxelDownSample: origin = min_bound - voxel/2, voxel
    index = floor((p - origin)/voxel), output = mean of the points in each voxel.
    open3d's output order is unordered_map order (not reproducible); this oracle
    returns voxels sorted by (ix, iy, iz).  Returns fp64 [m,3]."""
    p = np.asarray(points, np.float64)
    if len(p) == 0:
        return p.reshape(0, 3)
    origin = p.min(axis=0) - voxel_size * 0.5
    vi = np.floor((p - origin) / voxel_size).astype(np.int64)
    keys, inv = np.unique(vi, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    out = np.zeros((len(keys), 3))
    cnt = np.bincount(inv, minlength=len(keys)).astype(np.float64)
    for a in range(3):
        out[:, a] = np.bincount(inv, weights=p[:, a], minlength=len(keys))
    return out / cnt[:, None]


def icp_regression(target, source, voxel_size=5, threshold=100, use_kdtree=False):
    """open3d_utils.py:63-122 with the configuration the reference runs
    (global_regression=False, icp_point2point=True, icp_point2plane=False;
    main.py:177-179, create_labels.py:229-231).  The normals/FPFH that
    preprocess_point_cloud also computes (:23-32) feed only disabled branches.
    Returns (target_down, source_down, T)."""
    tgt = voxel_down_sample(target, voxel_size)
    src = voxel_down_sample(source, voxel_size)
    T = registration_icp_p2p(src, tgt, threshold, use_kdtree=use_kdtree)
    return tgt, src, T


# ----------------------------------------------------------------------------- outlier filters (open3d 0.9.0)
def _sq_dists(points, lo, hi):
    """[hi-lo, n] squared distances, fp64, ((dx^2 + dy^2) + dz^2) -- FLANN's L2 functor order."""
    q = points[lo:hi]
    d = None
    for a in range(3):
        diff = points[None, :, a] - q[:, None, a]
        sq = diff * diff
        d = sq if d is None else d + sq
    return d


def remove_radius_outlier(points, nb_points, radius, chunk=512):
    """pcd.remove_radius_outlier (open3d_utils.py:161, :205; Open3D v0.9.0 PointCloud::RemoveRadiusOutliers): keep point i
    iff more than `nb_points` points (itself included) lie strictly within `radius` (KDTreeFlann::SearchRadius ->
    FLANN radius search, d^2 < r^2).  Returns (kept points, kept indices)."""
    points = np.asarray(points, np.float64)
    keep = np.zeros(len(points), bool)
    r2 = radius * radius
    for s in range(0, len(points), chunk):
        keep[s:s + chunk] = (_sq_dists(points, s, min(len(points), s + chunk)) < r2).sum(axis=1) > nb_points
    idx = np.nonzero(keep)[0]
    return points[idx], idx


def compute_mahalanobis_distance(points):
    """pcd.compute_mahalanobis_distance (open3d_utils.py:163, :200, :207): mean / POPULATION covariance from the
    cumulants (PointCloud::ComputeMeanAndCovariance, sequential sums), then sqrt(p^T cov^-1 p)."""
    points = np.asarray(points, np.float64)
    cu = np.zeros(9)
    for p in points:                                        # sequential accumulation, as the C++ loop
        cu += (p[0], p[1], p[2], p[0] * p[0], p[0] * p[1], p[0] * p[2], p[1] * p[1], p[1] * p[2], p[2] * p[2])
    cu /= float(len(points))
    mean = cu[:3]
    cov = np.array([[cu[3] - cu[0] * cu[0], cu[4] - cu[0] * cu[1], cu[5] - cu[0] * cu[2]],
                    [cu[4] - cu[0] * cu[1], cu[6] - cu[1] * cu[1], cu[7] - cu[1] * cu[2]],
                    [cu[5] - cu[0] * cu[2], cu[7] - cu[1] * cu[2], cu[8] - cu[2] * cu[2]]])
    inv = np.linalg.inv(cov)
    d = points - mean
    return np.sqrt(np.einsum('ij,jk,ik->i', d, inv, d))


def remove_statistical_outlier(points, nb_neighbors, std_ratio, chunk=512):
    """pcd.remove_statistical_outlier (open3d_utils.py:165, :210; v0.9.0 PointCloud::RemoveStatisticalOutliers): per point
    the mean of the distances to its nb_neighbors nearest points (SearchKNN: itself included, ascending), cloud mean and
    Bessel-corrected std over the valid points accumulated SEQUENTIALLY (std::accumulate / inner_product), keep iff
    0 < avg < mean + std_ratio * std.  Returns (kept points, kept indices, avg distances, threshold)."""
    points = np.asarray(points, np.float64)
    n = len(points)
    avg = np.empty(n)
    k = min(nb_neighbors, n)
    for s in range(0, n, chunk):
        d2 = np.sort(_sq_dists(points, s, min(n, s + chunk)), axis=1)[:, :k]
        d = np.sqrt(d2)
        acc = np.zeros(len(d))
        for j in range(k):                                  # ascending order, as std::accumulate over the sorted result
            acc = acc + d[:, j]
        avg[s:s + chunk] = acc / float(k)
    valid = 0
    mean = 0.0
    for a in avg:
        if a > 0:
            mean += a; valid += 1
    mean /= valid
    sq = 0.0
    for a in avg:
        if a > 0:
            sq += (a - mean) * (a - mean)
    thr = mean + std_ratio * np.sqrt(sq / (valid - 1))
    idx = np.nonzero((avg > 0) & (avg < thr))[0]
    return points[idx], idx, avg, thr


def get_surface_filters(points, min_friends, min_dist, nb_neighbors, voxel_size):
    """The filter chain of get_surface (open3d_utils.py:198-211) on an already back-projected cloud:
    voxel grid -> radius outliers -> std of |Mahalanobis| -> statistical outliers with that std as ratio."""
    s = voxel_down_sample(points, voxel_size)
    s, _ = remove_radius_outlier(s, min_friends, min_dist)
    ratio = float(np.std(np.abs(compute_mahalanobis_distance(s))))
    s, _, _, _ = remove_statistical_outlier(s, nb_neighbors, ratio)
    return s
