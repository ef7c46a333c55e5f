"""GPU parity of the point-cloud outlier filters (open3d_utils.py:158-166, :198-211) against the numpy
restatement of the open3d 0.9.0 algorithms (oracle/icp.py).  Kept index sets must be identical (a point may differ only
when it sits within 1e-9 of a decision threshold); Mahalanobis distances and neighbour averages within 1e-9."""
import numpy as np
import pytest
import torch

from oracle import icp as oicp, geometry as og, synth

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _cloud(seed, n, outliers=30):
    rng = np.random.RandomState(seed)
    pts = synth.ellipsoid_cloud(rng, n) + rng.randn(n, 3) * 0.3
    far = rng.uniform(-120, 120, size=(outliers, 3))
    return np.concatenate([pts, far]).astype(np.float64)[rng.permutation(n + outliers)]


@pytest.mark.parametrize('n,nb,radius', [(1500, 5, 5.0), (3000, 5, 2.5), (700, 3, 8.0)])
def test_radius_outlier_vs_oracle(n, nb, radius):
    from autoposeestimation_b200 import ops
    pts = _cloud(n, n)
    off = np.array([0, len(pts)], np.int32)
    keep = ops.radius_outlier(_dev(pts), _dev(off), nb, radius, off).cpu().numpy().astype(bool)
    _, idx = oicp.remove_radius_outlier(pts, nb, radius)
    assert np.array_equal(np.nonzero(keep)[0], idx)
    assert 0 < keep.sum() < len(pts)                      # the far points are removed, the surface stays


def test_mahalanobis_vs_oracle():
    from autoposeestimation_b200 import ops
    pts = _cloud(3, 2500)
    off = np.array([0, len(pts)], np.int32)
    dist, std = ops.mahalanobis(_dev(pts), _dev(off))
    want = oicp.compute_mahalanobis_distance(pts)
    assert np.allclose(dist.cpu().numpy(), want, rtol=1e-9, atol=1e-9)
    assert abs(float(std[0]) - float(np.std(np.abs(want)))) < 1e-9


@pytest.mark.parametrize('n,k,ratio', [(1500, 20, 1.0), (2600, 20, 0.6), (900, 8, 2.0)])
def test_statistical_outlier_vs_oracle(n, k, ratio):
    from autoposeestimation_b200 import ops
    pts = _cloud(10 + n, n)
    off = np.array([0, len(pts)], np.int32)
    keep, avg, thr = ops.statistical_outlier(_dev(pts), _dev(off), k, ratio, off)
    _, idx, avg_ref, thr_ref = oicp.remove_statistical_outlier(pts, k, ratio)
    assert np.allclose(avg.cpu().numpy(), avg_ref, rtol=1e-12, atol=1e-12)
    assert abs(float(thr[0]) - thr_ref) < 1e-9
    want = np.zeros(len(pts), bool); want[idx] = True
    diff = np.nonzero(keep.cpu().numpy().astype(bool) != want)[0]
    assert all(abs(avg_ref[i] - thr_ref) < 1e-9 for i in diff)          # only points sitting on the threshold may differ
    assert len(diff) == 0 or len(diff) <= 1


def test_ragged_batch_compaction_and_edge_cases():
    """Three clouds of different sizes (one tiny, one empty) in one launch; compaction keeps the original order."""
    from autoposeestimation_b200 import ops
    clouds = [_cloud(1, 800), np.zeros((0, 3)), _cloud(2, 1300), np.array([[0.0, 0, 0], [0.1, 0, 0], [50.0, 0, 0]])]
    flat = np.concatenate(clouds)
    off = np.zeros(len(clouds) + 1, np.int32); off[1:] = np.cumsum([len(c) for c in clouds])
    keep = ops.radius_outlier(_dev(flat), _dev(off), 1, 5.0, off)
    out, cnt, index = ops.compact_points(_dev(flat), _dev(off), keep, want_index=True)
    cnt = cnt.cpu().numpy(); out = out.cpu().numpy(); index = index.cpu().numpy()
    for c, pts in enumerate(clouds):
        ref_pts, ref_idx = oicp.remove_radius_outlier(pts, 1, 5.0) if len(pts) else (pts, np.zeros(0, int))
        assert cnt[c] == len(ref_idx)
        assert np.array_equal(out[off[c]:off[c] + cnt[c]], ref_pts) and np.array_equal(index[off[c]:off[c] + cnt[c]], ref_idx)
    assert cnt[3] == 2                                                   # the two close points keep each other, the far one goes
    _, std = ops.mahalanobis(_dev(flat), _dev(off), want_dist=False)
    keep2, avg, thr = ops.statistical_outlier(_dev(flat), _dev(off), 20, std, off)      # per-cloud ratio from the device
    for c in (0, 2):
        pts = clouds[c]
        ratio = float(np.std(np.abs(oicp.compute_mahalanobis_distance(pts))))
        _, idx, _, _ = oicp.remove_statistical_outlier(pts, 20, ratio)
        assert np.array_equal(np.nonzero(keep2.cpu().numpy()[off[c]:off[c + 1]])[0], idx)


def test_pointcloud_methods_and_align_point_clouds():
    """The o3d-like methods the reference calls (open3d_utils.py:158-166) and align_point_clouds (:125-168)."""
    from autoposeestimation_b200.pc_reconstruction.open3d_utils import PointCloud, align_point_clouds
    pts = _cloud(7, 2000)
    pc = PointCloud(pts)
    f1, idx1 = pc.remove_radius_outlier(nb_points=5, radius=5)
    ref1, ridx1 = oicp.remove_radius_outlier(pts, 5, 5)
    assert np.array_equal(idx1.cpu().numpy(), ridx1) and np.array_equal(f1.numpy(), ref1)
    dist = f1.compute_mahalanobis_distance()
    assert np.allclose(dist, oicp.compute_mahalanobis_distance(ref1), rtol=1e-9, atol=1e-9)
    f2, idx2 = f1.remove_statistical_outlier(nb_neighbors=20, std_ratio=float(np.std(dist)))
    ref2, ridx2, _, _ = oicp.remove_statistical_outlier(ref1, 20, float(np.std(oicp.compute_mahalanobis_distance(ref1))))
    assert np.array_equal(idx2.cpu().numpy(), ridx2)
    # align: the same object seen in two rotation runs (second shifted by a small rigid motion)
    rng = np.random.RandomState(0)
    model = synth.ellipsoid_cloud(rng, 3000)
    R = synth.random_rotation(rng, 0.05); t = np.array([1.0, -40.0, 0.5])
    run2 = model @ R.T + t
    merged = align_point_clouds([PointCloud(model.copy()), PointCloud(run2)], 2, 6, 20, voxel_size=5, threshold=50)
    assert 0 < len(merged) <= 6000
    # reference flow with the oracle pieces
    tgt = model.copy(); src = run2.copy()
    diff = src.mean(0) - tgt.mean(0)
    if diff[1] > -30:
        src = src + np.array([0, -30 - diff[1], 0])
    td, sd, T = oicp.icp_regression(tgt, src, voxel_size=5, threshold=50)
    sd = sd @ T[:3, :3].T + T[:3, 3]
    want = oicp.voxel_down_sample(np.concatenate((sd, td)), 5)
    want, _ = oicp.remove_radius_outlier(want, 2, 6)
    ratio = float(np.std(oicp.compute_mahalanobis_distance(want)))
    want, _, _, _ = oicp.remove_statistical_outlier(want, 20, ratio)
    assert len(merged) == len(want) and np.allclose(merged.numpy(), want, atol=1e-6)


def test_get_surfaces_batch_and_reconstruct_run():
    """create_pointcloud.py:232-317: all views through batched launches == per-view get_surface; the sequential
    register-and-merge loop == the same flow built from the oracle pieces."""
    from autoposeestimation_b200.pc_reconstruction.open3d_utils import get_surface
    from autoposeestimation_b200.pc_reconstruction.create_pointcloud import get_surfaces_batch, reconstruct_run, rotate_about_center
    seeds = [20, 21, 22]
    frames = [synth.render_ellipsoid_frame(s, max_rot_deg=3.0, max_trans_mm=2.0) for s in seeds]
    lab = _dev(np.stack([f['label'] for f in frames])); dep = _dev(np.stack([f['depth'] for f in frames]).view(np.int16))
    intr = frames[0]['intr']
    cam = _dev(np.tile(np.array([[intr['ppx'], intr['ppy'], intr['fx'], intr['fy']]]), (3, 1)))
    r2c = _dev(np.stack([f['robot2cam'] for f in frames]))
    # an empty view in the middle must come back as an empty cloud
    lab[1].zero_()
    surfs = get_surfaces_batch(lab, dep, cam, r2c, 5, 5.0, 20, 2.0)
    assert len(surfs) == 3 and len(surfs[1]) == 0
    for v in (0, 2):
        single = get_surface(frames[v]['label'], frames[v]['depth'].astype(np.float64), intr, frames[v]['robot2cam'], 5, 5.0, 20, voxel_size=2.0)
        assert len(single) == len(surfs[v]) and np.array_equal(single.numpy(), surfs[v].numpy())
        pts, _ = og.surface_backproject(frames[v]['label'], frames[v]['depth'].astype(np.float64), intr, frames[v]['robot2cam'])
        want = oicp.get_surface_filters(pts, 5, 5.0, 20, 2.0)
        assert len(want) == len(surfs[v]) and np.allclose(surfs[v].numpy(), want, atol=1e-9)
    # sequential loop: the same physical surface seen twice with a small rigid offset -> register, merge, voxel
    a = surfs[0].numpy()
    R = synth.random_rotation(np.random.RandomState(3), 0.03); t = np.array([1.5, -1.0, 0.8])
    from autoposeestimation_b200.pc_reconstruction.open3d_utils import PointCloud
    b = PointCloud((a - a.mean(0)) @ R.T + a.mean(0) + t)
    cloud = reconstruct_run([surfs[0], surfs[1], b], voxel_size=2.0, threshold=10.0)
    td, sd, T = oicp.icp_regression(a, b.numpy(), voxel_size=2.0, threshold=10.0)
    want = oicp.voxel_down_sample(np.concatenate((sd @ T[:3, :3].T + T[:3, 3], td)), 2.0)
    assert len(cloud) == len(want) and np.allclose(cloud.numpy(), want, atol=1e-6)
    # the device-resident loop (csrc/reconstruct.cu: one call, sizes on the device) against the view-by-view loop, on a
    # longer sequence with an empty view in front and in the middle
    views = [surfs[1], surfs[0], b, surfs[1]]
    rng = np.random.RandomState(8)
    for k in range(4):
        Rk = synth.random_rotation(rng, 0.02); tk = rng.uniform(-1.5, 1.5, size=3)
        sub = a[rng.rand(len(a)) < 0.8]                                           # partial overlap
        views.append(PointCloud((sub - a.mean(0)) @ Rk.T + a.mean(0) + tk))
    dev_cloud = reconstruct_run(views, voxel_size=2.0, threshold=10.0)
    host_cloud = reconstruct_run(views, voxel_size=2.0, threshold=10.0, device_loop=False)
    assert len(dev_cloud) == len(host_cloud) and np.allclose(dev_cloud.numpy(), host_cloud.numpy(), atol=1e-6)
    assert len(cloud) == len(reconstruct_run([surfs[0], surfs[1], b], voxel_size=2.0, threshold=10.0, device_loop=False))
    assert len(reconstruct_run([surfs[1], surfs[1]], 2.0, 10.0)) == 0 and len(reconstruct_run([surfs[1], b], 2.0, 10.0)) == len(b)
    c0 = cloud.numpy().mean(0)
    rot = rotate_about_center(cloud, R)
    assert np.allclose(rot.numpy().mean(0), c0, atol=1e-9)
