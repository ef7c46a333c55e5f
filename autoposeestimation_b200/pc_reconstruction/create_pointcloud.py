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
        out, vc = ops.voxel_down_sample(flat, torch.from_numpy(off).to(dev), float(voxel_size), offset_host=off)
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


def reconstruct_run(surfaces, voxel_size, threshold, global_regression=False, icp_point2point=True, icp_point2plane=False,
                    device_loop=True):
    """create_pointcloud.py:286-312: first non-empty surface starts the cloud; every further one is registered to it
    (`icp_regression`), transformed, concatenated IN FRONT of the target and voxel-down-sampled."""
    live = [s for s in surfaces if len(s) > 0]
    if device_loop and len(live) >= 2 and icp_point2point and not global_regression and not icp_point2plane:
        # the whole loop in one C call, sizes on the device, ONE synchronisation at the end (csrc/reconstruct.cu)
        off = np.zeros(len(live) + 1, np.int32)
        off[1:] = np.cumsum([len(s) for s in live])
        out, cnt, status = ops.reconstruct_run(torch.cat([s.points for s in live]), off, voxel_size, threshold)
        c, st = (int(v) for v in torch.cat((cnt, status)).cpu())
        if st == 0:
            return PointCloud(out[:c].clone())
        # an intermediate cloud outgrew the batched voxel kernel: view by view below (large clouds take the global-memory path)
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


# ------------------------------------------------------------------------------------------------------------------
# Entry point of main.py option 4 with the reference's signature (create_pointcloud.py:181-378)
def bbox_center(cloud):
    """`utils.get_my_source_center` (open3d_utils.py:273-290): centre of the axis-aligned bounding box."""
    p = cloud.points
    return ((p.min(dim=0).values + p.max(dim=0).values) / 2).cpu().numpy()


def get_view_distribution(data_path, d, n, n_viewpoints, plot=False, l_arrow=30, reference_point=np.array([0, 0, 0])):
    """create_pointcloud.py:46-179 (view selection; the 3-D plot is outside the graft): camera positions of the n frames
    of rotation run `d` -> voxel grid whose cell size is grown / shrunk by 1 mm until n_viewpoints cells are occupied ->
    the frame nearest to each cell mean -> greedy nearest-neighbour chain starting at the view closest to the robot origin.
    Returns frame indices in visiting order.  (When the grid cannot hit n_viewpoints exactly the reference draws a random
    subset of the cells, :99-104; the draw comes from np.random here as well, over cells in sorted instead of hash order.)"""
    import json
    import os
    pos = []
    for idx in range(n):
        with open(os.path.join(data_path, d, '{:06d}.meta.json'.format(idx))) as f:
            meta = json.load(f)
        r2c = np.dot(np.array(meta.get('robot2endEff_tf'), np.float64).reshape(4, 4),
                     np.array(meta.get('hand_eye_calibration'), np.float64).reshape(4, 4))
        pos.append(r2c[:3, 3])
    pos = np.array(pos)
    diff = np.linalg.norm(pos[:, None, :] - pos[None, :, :], axis=2)
    voxel = int(diff[~np.eye(len(pos), dtype=bool)].min()) if len(pos) > 1 else 1
    cams = PointCloud(pos)
    while True:
        cells = cams.voxel_down_sample(max(voxel, 1e-9)).numpy() if voxel > 0 else pos
        if len(cells) == n_viewpoints:
            centres = cells
            break
        if len(cells) < n_viewpoints:
            voxel -= 1
            cells = cams.voxel_down_sample(voxel).numpy() if voxel > 0 else pos
            centres = cells[np.random.choice(np.arange(len(cells)), replace=False, size=n_viewpoints)]
            break
        voxel += 1
    nearest = [int(np.argmin(np.linalg.norm(pos - c, axis=1))) for c in centres]
    chosen = pos[nearest]
    order = [int(np.argmin(np.linalg.norm(chosen, axis=1)))]
    while len(order) != n_viewpoints:
        rest = [j for j in range(len(chosen)) if j not in order]
        order.append(rest[int(np.argmin([np.linalg.norm(chosen[j] - chosen[order[-1]]) for j in rest]))])
    return np.array(nearest)[order]


def load_point_cloud(object_name, save_dir, root, reference_point=np.array([0, 0, 0]), mode='gen', n_viewpoints=10, min_friends=10,
                     voxel_size=5, voxel_size_out=10, threshold=50, min_dist=10, nb_neighbors=5, l_arrow=30,
                     global_regression=False, icp_point2point=True, icp_point2plane=True, plot=False):
    """Drop-in for create_pointcloud.py:181-378 (same arguments, same files written, returns the aligned cloud).

    Per rotation run under `label_generator/data/<object>/`: select n_viewpoints views, decode them into pinned buffers
    (formats.FrameBatchLoader), run `get_surface` for ALL of them in one pass of batched launches, register-and-merge them
    sequentially (`reconstruct_run`), rotate by the run's object pose, write `<run>.pcd/.ply`; then align the runs
    (`align_point_clouds`) and write `<object>_out`, the down-sampled centred `<object>` cloud and the `.xyz` model the
    pose networks train on (>= 1000 points, :363-376).  As in the reference, `icp_point2plane=True` is not grafted and
    raises inside icp_regression (main.py:177-179 passes False); plots are outside the graft."""
    import os
    from .. import formats
    from .open3d_utils import align_point_clouds
    label_root = os.path.join(root, 'label_generator/data', object_name)
    runs = [d for d in os.listdir(label_root) if d != 'extra']
    if not runs:
        raise ValueError('no labels obtained yet')
    data_root = os.path.join(root, 'data_generation/data', object_name)
    out_dir = os.path.join(save_dir, object_name)
    os.makedirs(out_dir, exist_ok=True)
    n = len([f for f in os.listdir(os.path.join(label_root, runs[0])) if '.{}.label.png'.format(mode) in f])
    run_clouds = []
    for d in runs:
        views = get_view_distribution(data_root, d, n, n_viewpoints, plot=False, l_arrow=l_arrow, reference_point=reference_point)
        loader = formats.FrameBatchLoader(os.path.join(data_root, d), os.path.join(label_root, d), mode=mode)
        batch = loader.to_device(loader.load([int(v) for v in views]))
        surfaces = get_surfaces_batch(batch['label'], batch['depth'], batch['cam'], batch['robot2cam'], min_friends, min_dist,
                                      nb_neighbors, voxel_size)
        cloud = reconstruct_run(surfaces, voxel_size, threshold, global_regression=global_regression,
                                icp_point2point=icp_point2point, icp_point2plane=icp_point2plane)
        rot = batch['meta'][-1]['object_rotation']                 # the last view's object_pose, as :246-247 leaves it
        if len(cloud):
            rotate_about_center(cloud, rot)
        formats.write_pcd(os.path.join(out_dir, '{}.pcd'.format(d)), cloud.numpy())
        formats.write_ply(os.path.join(out_dir, '{}.ply'.format(d)), cloud.numpy())
        run_clouds.append(PointCloud(cloud.points.clone()))
    cloud = align_point_clouds(run_clouds, min_friends=min_friends, min_dist=min_dist, nb_neighbors=nb_neighbors, plot=False,
                               global_regression=global_regression, icp_point2point=icp_point2point,
                               icp_point2plane=icp_point2plane, voxel_size=voxel_size, threshold=threshold)
    formats.write_pcd(os.path.join(out_dir, '{}_out.pcd'.format(object_name)), cloud.numpy())
    formats.write_ply(os.path.join(out_dir, '{}_out.ply'.format(object_name)), cloud.numpy())
    down = cloud.voxel_down_sample(voxel_size_out)
    down.translate(-bbox_center(down))
    formats.write_pcd(os.path.join(out_dir, '{}.pcd'.format(object_name)), down.numpy())
    formats.write_ply(os.path.join(out_dir, '{}.ply'.format(object_name)), down.numpy())
    big = PointCloud(cloud.points.clone())
    big.translate(-bbox_center(big))
    v = voxel_size                                             # coarsen in 0.1 mm steps while at least 1000 points remain (:363-370)
    while True:
        v += 0.1
        if len(big.voxel_down_sample(v)) < 1000:
            big = big.voxel_down_sample(v - 0.1)
            break
    formats.write_xyz(os.path.join(out_dir, '{}.xyz'.format(object_name)), big.numpy())
    return cloud
