"""Generate tests/golden/geometry_ref.npz by running the REFERENCE's own geometry code on seeded synthetic frames.

Run here (dev container) only:  python -m oracle.gen_golden_geometry
TEST INFRASTRUCTURE (see oracle/__init__.py).  No reference source is copied: the modules are imported unmodified
from /root/reference; only their third-party imports that are absent from this image (open3d, matplotlib, mathutils,
transforms3d) are replaced by empty stand-ins, none of which does arithmetic on the pinned values.

Reference code exercised:
  DenseFusion/datasets/myDatasetAugmented/dataset.py
      get_bbox (:342-380)                                   -> row a1
      PoseDataset.__getitem__ (:157-318), mode 'test', add_noise=False: mask, bbox, `choose` with
      np.random.shuffle of the 0/1 vector or np.pad(...,'wrap'), fp32 back-projection (:236-275)   -> rows a2, a3
      (this is the same code as pipeline/utils.py:524-553, which lives inside full_prediction)
      PoseDataset.__init__ (:119-141): the .xyz model parser (= pipeline/utils.py:667-684)          -> row (f)-4
  pc_reconstruction/open3d_utils.py
      get_surface (:171-192): the per-pixel fp64 loop; the open3d calls that follow it (:194-212) are
      stand-ins that return the cloud unchanged, so the fixture holds the raw back-projected points   -> row a4

The dataset class reads files, so a small dataset tree is written to a temporary directory in the reference's own
on-disk formats (16-bit depth PNG, 8-bit label PNG, meta.json, .xyz, list files) and deleted afterwards.
"""
import json
import os
import random
import shutil
import sys
import tempfile
import types
import warnings

import numpy as np

REF = os.environ.get('APE_REFERENCE', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')


class _CapturedCloud:
    """Stand-in for o3d.geometry.PointCloud: holds what get_surface assigns; every filter is the identity."""

    def __init__(self):
        self.points = np.zeros((0, 3))

    def voxel_down_sample(self, voxel_size):
        return self

    def compute_mahalanobis_distance(self):
        return np.zeros(max(len(self.points), 1))

    def remove_radius_outlier(self, nb_points, radius):
        return self, []

    def remove_statistical_outlier(self, nb_neighbors, std_ratio):
        return self, []


def _stub_missing_imports():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    o3d = mod('open3d')
    o3d.geometry = mod('open3d.geometry', PointCloud=_CapturedCloud)
    o3d.utility = mod('open3d.utility', Vector3dVector=lambda a: np.array(a))
    mp = mod('matplotlib')
    mp.pyplot = mod('matplotlib.pyplot')
    mu = mod('mathutils')
    mu.geometry = mod('mathutils.geometry', intersect_line_line=None)
    t3 = mod('transforms3d')
    t3.euler = mod('transforms3d.euler')
    if not hasattr(np, 'float'):
        np.float = float          # removed numpy alias the reference still uses (create_pointcloud.py:250)


def _import_reference():
    sys.path.insert(0, REF)
    warnings.filterwarnings('ignore')
    _stub_missing_imports()
    import DenseFusion.datasets.myDatasetAugmented.dataset as dataset
    import pc_reconstruction.open3d_utils as o3u
    return dataset, o3u


def _frames():
    """Three 480x640 frames: (0) large object -> shuffle branch; (1) object touching the top-left corner -> bbox
    clamping, shuffle branch; (2) tiny object -> fewer candidates than num_pt -> 'wrap' branch."""
    from oracle import synth
    out = []
    f0 = synth.render_ellipsoid_frame(11)
    out.append((f0['depth'], f0['label']))
    f1 = synth.render_ellipsoid_frame(12)
    rr, cc = np.nonzero(f1['label'])
    dr, dc = rr.min() + 3, cc.min() + 3           # shift the object so that it is cut by the image corner
    out.append((np.roll(f1['depth'], (-dr, -dc), (0, 1)), np.roll(f1['label'], (-dr, -dc), (0, 1))))
    out[1][1][-dr:, :] = 0; out[1][1][:, -dc:] = 0
    f2 = synth.render_ellipsoid_frame(13)
    lab = f2['label'].copy()
    rr, cc = np.nonzero(lab)
    keep = np.zeros_like(lab)
    r0, c0 = int(rr.mean()), int(cc.mean())
    keep[r0 - 7:r0 + 8, c0 - 9:c0 + 10] = lab[r0 - 7:r0 + 8, c0 - 9:c0 + 10]
    out.append((f2['depth'], keep))
    # depth outside the object never reaches a pinned value (mask = label & depth); flatten it so the fixture stays small
    out = [(np.where(lab != 0, dep, 1000).astype(np.uint16), lab) for dep, lab in out]
    return out, f0['intr'], f0['robot2cam']


def _write_tree(root, frames, intr, depth_scale, model_mm):
    from PIL import Image
    cls, run, ds = 'ellipsoid', 'run0', 'golden'
    j = os.path.join
    for d in (j(root, 'data_generation/data', cls, run), j(root, 'label_generator/data', cls, run),
              j(root, 'label_generator/data_sets/pose_estimation', ds), j(root, 'pc_reconstruction/data', cls)):
        os.makedirs(d)
    rng = np.random.RandomState(5)
    lines = []
    poses = []
    for i, (depth, label) in enumerate(frames):
        fid = '%s/%s/%06d' % (cls, run, i)
        lines.append(fid)
        Image.fromarray(rng.randint(0, 255, (480, 640, 3)).astype(np.uint8)).save(j(root, 'data_generation/data', fid + '.color.png'))
        Image.fromarray(depth.astype(np.uint16)).save(j(root, 'data_generation/data', fid + '.depth.png'))
        Image.fromarray(label.astype(np.uint8)).save(j(root, 'label_generator/data', fid + '.new_pred.label.png'))
        with open(j(root, 'data_generation/data', fid + '.meta.json'), 'w') as f:
            json.dump({'intr': intr, 'depth_scale': depth_scale, 'symmetric': False, 'view_point_id': i}, f)
        cam2robot = np.eye(4); robot2object = np.eye(4)
        ang = 0.3 + 0.2 * i
        robot2object[:3, :3] = [[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]]
        robot2object[:3, 3] = [10.0 * i, -20.0, 600.0]
        poses.append(robot2object)
        with open(j(root, 'label_generator/data', fid + '.meta.json'), 'w') as f:
            json.dump({'cls_name': cls, 'cam2robot': cam2robot.reshape(-1).tolist(), 'robot2object': robot2object.reshape(-1).tolist()}, f)
    with open(j(root, 'label_generator/data_sets/pose_estimation', ds, 'test_data_list.txt'), 'w') as f:
        f.write(''.join(l + '\n' for l in lines))
    with open(j(root, 'label_generator/data_sets/pose_estimation', ds, 'classes.txt'), 'w') as f:
        f.write(cls + '\n')
    with open(j(root, 'pc_reconstruction/data', cls, cls + '.xyz'), 'w') as f:      # create_pointcloud.py:373-376
        for p in model_mm:
            f.write('%s\n' % p)
    return ds


def main():
    from oracle import synth
    dataset, o3u = _import_reference()
    frames, intr, robot2cam = _frames()
    depth_scale = 0.0010000000474974513          # RealSense D4xx value, written by getData.py into meta.json
    model_mm = synth.ellipsoid_cloud(np.random.RandomState(3), 1200)
    num_pt = 500
    seeds = [21, 22, 23]
    root = tempfile.mkdtemp(prefix='ape_golden_')
    try:
        ds_name = _write_tree(root, frames, intr, depth_scale, model_mm)
        ds = dataset.PoseDataset('test', num_pt, False, 0.0, False, ds_name, root)
        xyz_text = open(os.path.join(root, 'pc_reconstruction/data/ellipsoid/ellipsoid.xyz')).read()
        parsed_model = np.array(ds.cld[0], np.float64)            # the reference's own .xyz parser (dataset.py:119-141)
        clouds, chooses, bboxes, targets, models = [], [], [], [], []
        for i, s in enumerate(seeds):
            np.random.seed(s); random.seed(s)
            cloud, choose, img, target, model_points, obj, _, _ = ds[i]
            depth, label = frames[i]
            bboxes.append(dataset.get_bbox(label == 255))
            clouds.append(cloud.numpy()); chooses.append(choose.numpy()[0].astype(np.int64))
            targets.append(target.numpy()); models.append(model_points.numpy())
            assert tuple(img.shape[1:]) == (bboxes[-1][1] - bboxes[-1][0], bboxes[-1][3] - bboxes[-1][2])
    finally:
        shutil.rmtree(root)

    # get_bbox alone on random rectangles / blobs, incl. every clamp branch and exact multiples of 40
    rng = np.random.RandomState(7)
    rects, rect_bbox = [], []
    for k in range(64):
        h = int(rng.choice([1, 39, 40, 41, 80, 119, 120, 200, 241, 479, 480])) if k < 32 else int(rng.randint(1, 481))
        w = int(rng.choice([1, 39, 40, 41, 160, 161, 320, 600, 639, 640])) if k < 32 else int(rng.randint(1, 641))
        r0 = int(rng.randint(0, 480 - h + 1)); c0 = int(rng.randint(0, 640 - w + 1))
        m = np.zeros((480, 640), bool); m[r0:r0 + h, c0:c0 + w] = True
        rects.append((r0, r0 + h, c0, c0 + w)); rect_bbox.append(dataset.get_bbox(m))

    # get_surface: the per-pixel fp64 loop (filters are identity stand-ins)
    surf_pts = []
    for depth, label in frames:
        pc = o3u.get_surface(label, np.array(depth, dtype=np.float64), intr, robot2cam, 5, 2.0, 30, 2.0)
        surf_pts.append(np.asarray(pc.points, dtype=np.float64))

    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, 'geometry_ref.npz')
    np.savez_compressed(
        path, depth=np.stack([f[0] for f in frames]).astype(np.uint16), label=np.stack([f[1] for f in frames]).astype(np.uint8),
        intr=np.array([intr['ppx'], intr['ppy'], intr['fx'], intr['fy']], np.float64), depth_scale=np.float64(depth_scale),
        robot2cam=np.asarray(robot2cam, np.float64), num_pt=np.int64(num_pt), seeds=np.array(seeds, np.int64),
        bbox=np.array(bboxes, np.int64), choose=np.stack(chooses), cloud=np.stack(clouds),
        rects=np.array(rects, np.int64), rect_bbox=np.array(rect_bbox, np.int64),
        surf_n=np.array([len(p) for p in surf_pts], np.int64), surf_pts=np.concatenate(surf_pts),
        xyz_text=np.array(xyz_text), xyz_written=np.asarray(model_mm, np.float64), xyz_parsed_m=parsed_model)
    print('wrote', path, os.path.getsize(path), 'bytes;', 'candidates per frame:',
          [int(((f[1] == 255) & (f[0] != 0)).sum()) for f in frames], 'bbox', bboxes)


if __name__ == '__main__':
    main()
