"""On-disk formats on either side of the hot path and a batched frame loader that feeds the
label-path kernels from pinned host memory.  Host-side Python only (file I/O is not GPU work); every reader / writer
mirrors a specific piece of the reference so that files written by one side are read by the other:

  per-frame meta   `<id>.meta.json`   writer data_generation/getData.py:177-221, readers create_pointcloud.py:237-247,
                                      create_labels.py:396-401, main.py:507-515
  pose label       `<id>.meta.json`   writer label_generator/create_labels.py:405-429
  depth / label    `<id>.depth.png` (16 bit), `<id>.<mode>.label.png` (8 bit)   create_pointcloud.py:249-253
  model cloud      `<object>.xyz`     writer create_pointcloud.py:373-376, parser pipeline/utils.py:667-684
"""
import json
import os

import numpy as np


# ----------------------------------------------------------------------------------------------------- .xyz
def write_xyz(path, points):
    """create_pointcloud.py:373-376: one `"%s\\n" % row` per point, i.e. numpy's str() of a 3-vector
    (`[x y z]`, 8 significant digits, padded columns)."""
    with open(path, 'w') as f:
        for item in np.asarray(points):
            f.write("%s\n" % item)


def read_xyz(path, to_meter=True, exact=False):
    """pipeline/utils.py:667-684.  The reference strips `[` and `]\\n` with `readline()[1:-2]` and then drops ONE MORE
    character with `[:-1]` before splitting -- the last character of the z column is lost whenever numpy did not pad
    it (a reference quirk that the drop-in keeps so that both sides see the same model; `exact=True` parses the full
    number instead).  Values are divided by 1000 when `to_meter` (:678-679)."""
    pts = []
    with open(path) as f:
        while True:
            line = f.readline()[1:-2]
            if not line:
                break
            if not exact:
                line = line[:-1]
            xyz = [float(tok) / 1000 if to_meter else float(tok) for tok in line.split(' ') if tok != '']
            pts.append([xyz[0], xyz[1], xyz[2]])
    return np.array(pts)


# ----------------------------------------------------------------------------------------------------- meta.json
def frame_meta(joints, pose, object_tf, robot2endeff_tf, intr, depth_scale, symmetric, hand_eye_calibration, view_point_id):
    """The dict data_generation/getData.py:177-221 writes (4x4 matrices flattened row-major to 16 floats)."""
    flat = lambda m: [float(v) for v in np.asarray(m, np.float64).reshape(-1)]
    return {'joints': list(joints), 'pose': pose, 'object_pose': flat(object_tf), 'robot2endEff_tf': flat(robot2endeff_tf),
            'intr': {k: intr[k] for k in ('width', 'height', 'ppx', 'ppy', 'fx', 'fy', 'coeffs') if k in intr},
            'depth_scale': depth_scale, 'symmetric': symmetric, 'hand_eye_calibration': hand_eye_calibration,
            'view_point_id': view_point_id}


def write_json(path, obj):
    with open(path, 'w') as f:
        json.dump(obj, f)


def load_frame_meta(path):
    """What the label path needs from a frame's meta.json (create_pointcloud.py:237-247):
    intr, depth_scale, robot2Cam = robot2endEff . handEye (4x4 fp64, millimetres), object rotation 3x3."""
    with open(path) as f:
        meta = json.load(f)
    hand_eye = np.array(meta.get('hand_eye_calibration'), np.float64).reshape(4, 4)
    robot2endeff = np.array(meta.get('robot2endEff_tf'), np.float64).reshape(4, 4)
    obj = meta.get('object_pose')
    return {'intr': meta.get('intr'), 'depth_scale': meta.get('depth_scale'), 'hand_eye': hand_eye, 'robot2endEff': robot2endeff,
            'robot2Cam': np.dot(robot2endeff, hand_eye),
            'object_rotation': None if obj is None else np.array(obj, np.float64).reshape(4, 4)[:3, :3],
            'symmetric': meta.get('symmetric'), 'view_point_id': meta.get('view_point_id'), 'raw': meta}


def pose_label(hand_eye, robot2endeff, pc_rotation, pc_position, object_name):
    """create_labels.py:405-420: cam2robot = inv(handEye) . inv(robot2endEff); cam2object = cam2robot . robot2object."""
    robot2object = np.identity(4)
    robot2object[:3, :3] = pc_rotation
    robot2object[:3, 3] = pc_position
    cam2robot = np.dot(np.linalg.inv(hand_eye), np.linalg.inv(robot2endeff))
    cam2object = np.dot(cam2robot, robot2object)
    return {'position': list(cam2object[:3, 3]), 'rotation': list(cam2object[:3, :3].flatten()), 'cls_name': object_name,
            'cam2robot': list(cam2robot.flatten()), 'robot2object': list(robot2object.flatten())}


# ----------------------------------------------------------------------------------------------------- images
def _imread(path):
    import cv2
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if img is None:
        raise ValueError('cannot read image %s' % path)
    return img


def load_depth_png(path):
    """16-bit depth in raw sensor units (`np.array(Image.open(f))`, create_pointcloud.py:249-250) -> uint16 [H,W]."""
    d = _imread(path)
    if d.dtype != np.uint16 or d.ndim != 2:
        raise ValueError('%s: expected a single-channel 16-bit depth image, got %s %s' % (path, d.dtype, d.shape))
    return d


def load_label_png(path):
    """8-bit label (create_pointcloud.py:252-253) -> uint8 [H,W] (first channel of a colour image, as PIL -> uint8 gives
    for the single-channel files the label generator writes)."""
    lab = _imread(path)
    if lab.ndim == 3:
        lab = lab[:, :, 0]
    return lab.astype(np.uint8, copy=False)


def save_png(path, array):
    import cv2
    if not cv2.imwrite(path, array):
        raise ValueError('cannot write image %s' % path)


# ----------------------------------------------------------------------------------------------------- batched loader
class FrameBatchLoader:
    """Decodes the frames of one object run into PINNED host buffers shaped for `ops.surface_backproject`:
    depth [F,H,W] int16 storage of the uint16 values, label [F,H,W] uint8, cam [F,4] fp64 (ppx, ppy, fx, fy),
    robot2cam [F,4,4] fp64.  `to_device()` issues the four async copies on the current stream.
    File naming as the reference: `{:06d}.depth.png`, `{:06d}.meta.json` under `data_dir`,
    `{:06d}.{mode}.label.png` under `label_dir` (create_pointcloud.py:237-253)."""

    def __init__(self, data_dir, label_dir, mode='new_pred', height=480, width=640, pin=True, workers=None):
        self.data_dir, self.label_dir, self.mode, self.h, self.w, self.pin = data_dir, label_dir, mode, height, width, pin
        self.workers = workers if workers is not None else min(16, os.cpu_count() or 1)

    def load(self, indices):
        import torch
        F = len(indices)
        mk = lambda shape, dt: (torch.empty(shape, dtype=dt).pin_memory() if self.pin and torch.cuda.is_available()
                                else torch.empty(shape, dtype=dt))
        depth = mk((F, self.h, self.w), torch.int16); label = mk((F, self.h, self.w), torch.uint8)
        cam = mk((F, 4), torch.float64); r2c = mk((F, 4, 4), torch.float64)
        def one(k):
            idx = indices[k]
            m = load_frame_meta(os.path.join(self.data_dir, '{:06d}.meta.json'.format(idx)))
            d = load_depth_png(os.path.join(self.data_dir, '{:06d}.depth.png'.format(idx)))
            lab = load_label_png(os.path.join(self.label_dir, '{:06d}.{}.label.png'.format(idx, self.mode)))
            if d.shape != (self.h, self.w) or lab.shape != (self.h, self.w):
                raise ValueError('frame %d: expected %dx%d images, got depth %s label %s' % (idx, self.h, self.w, d.shape, lab.shape))
            depth[k].copy_(torch.from_numpy(d.view(np.int16)))
            label[k].copy_(torch.from_numpy(lab))
            intr = m['intr']
            cam[k] = torch.tensor([intr['ppx'], intr['ppy'], intr['fx'], intr['fy']], dtype=torch.float64)
            r2c[k] = torch.from_numpy(m['robot2Cam'])
            return m
        # PNG decoding releases the GIL (OpenCV): decode the frames of a run on a thread pool straight into the pinned buffers
        workers = min(self.workers, F) if F else 1
        if workers > 1:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=workers) as ex:
                metas = list(ex.map(one, range(F)))
        else:
            metas = [one(k) for k in range(F)]
        return {'depth': depth, 'label': label, 'cam': cam, 'robot2cam': r2c, 'meta': metas, 'indices': list(indices)}

    @staticmethod
    def to_device(batch, device=None):
        import torch
        dev = device or torch.device('cuda', torch.cuda.current_device())
        return {k: (v.to(dev, non_blocking=True) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}


# ----------------------------------------------------------------------------------------------------- .ply / .pcd clouds
def write_ply(path, points):
    """`o3d.io.write_point_cloud(path.ply, cloud)` (create_pointcloud.py:324-354): binary little-endian PLY, one vertex
    element with double x, y, z (what open3d writes for a cloud without colours / normals)."""
    pts = np.ascontiguousarray(np.asarray(points, np.float64).reshape(-1, 3))
    header = ('ply\nformat binary_little_endian 1.0\ncomment written by autoposeestimation_b200\nelement vertex %d\n'
              'property double x\nproperty double y\nproperty double z\nend_header\n' % len(pts))
    with open(path, 'wb') as f:
        f.write(header.encode('ascii'))
        f.write(pts.astype('<f8').tobytes())


_PLY_TYPES = {'char': 'i1', 'uchar': 'u1', 'short': 'i2', 'ushort': 'u2', 'int': 'i4', 'uint': 'u4', 'float': 'f4', 'double': 'f8',
              'int8': 'i1', 'uint8': 'u1', 'int16': 'i2', 'uint16': 'u2', 'int32': 'i4', 'uint32': 'u4', 'float32': 'f4', 'float64': 'f8'}


def read_ply(path):
    """`o3d.io.read_point_cloud(path.ply)` (create_labels.py:331, :346-347): the x, y, z of the vertex element as fp64 [n,3];
    ascii and binary (little / big endian) files, extra vertex properties (colours, normals) are skipped."""
    with open(path, 'rb') as f:
        if f.readline().strip() != b'ply':
            raise ValueError('%s: not a PLY file' % path)
        fmt, n, props, in_vertex = None, 0, [], False
        while True:
            line = f.readline()
            if not line:
                raise ValueError('%s: truncated PLY header' % path)
            tok = line.decode('ascii', 'replace').split()
            if not tok:
                continue
            if tok[0] == 'format':
                fmt = tok[1]
            elif tok[0] == 'element':
                in_vertex = tok[1] == 'vertex'
                if in_vertex:
                    n = int(tok[2])
            elif tok[0] == 'property' and in_vertex:
                if tok[1] == 'list':
                    raise ValueError('%s: list property in the vertex element' % path)
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == 'end_header':
                break
        names = [p[0] for p in props]
        if not all(a in names for a in 'xyz'):
            raise ValueError('%s: vertex element without x, y, z' % path)
        if fmt == 'ascii':
            rows = np.loadtxt(f, max_rows=n, ndmin=2) if n else np.zeros((0, len(props)))
            return np.ascontiguousarray(rows[:, [names.index(a) for a in 'xyz']], dtype=np.float64)
        order = '<' if fmt == 'binary_little_endian' else '>'
        rec = np.dtype([(nm, order + ty) for nm, ty in props])
        data = np.frombuffer(f.read(n * rec.itemsize), dtype=rec, count=n)
        return np.stack([data[a].astype(np.float64) for a in 'xyz'], axis=1)


def write_pcd(path, points):
    """`o3d.io.write_point_cloud(path.pcd, cloud)`: PCD v0.7, float32 x y z, binary (written for the user's viewers only;
    nothing on the path reads it back)."""
    pts = np.ascontiguousarray(np.asarray(points, np.float64).reshape(-1, 3).astype('<f4'))
    n = len(pts)
    header = ('# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n'
              'WIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA binary\n' % (n, n))
    with open(path, 'wb') as f:
        f.write(header.encode('ascii'))
        f.write(pts.tobytes())


# ----------------------------------------------------------------------------------------------------- Euler angles
_EPS4 = np.finfo(float).eps * 4.0


def mat2euler(M):
    """transforms3d.euler.mat2euler(M) with its default axes 'sxyz' (create_labels.py:344, :363-366): static x-y-z angles
    of a rotation matrix (the closed form of Gohlke's euler_from_matrix, DenseFusion/lib/transformations.py:1057-1112)."""
    M = np.asarray(M, np.float64)[:3, :3]
    cy = np.sqrt(M[0, 0] * M[0, 0] + M[1, 0] * M[1, 0])
    if cy > _EPS4:
        return (float(np.arctan2(M[2, 1], M[2, 2])), float(np.arctan2(-M[2, 0], cy)), float(np.arctan2(M[1, 0], M[0, 0])))
    return (float(np.arctan2(-M[1, 2], M[1, 1])), float(np.arctan2(-M[2, 0], cy)), 0.0)


def euler2mat(ax, ay, az):
    """transforms3d.euler.euler2mat(ai, aj, ak) with axes 'sxyz' (create_labels.py:372): R = Rz(az) Ry(ay) Rx(ax)."""
    sx, sy, sz = np.sin(ax), np.sin(ay), np.sin(az)
    cx, cy, cz = np.cos(ax), np.cos(ay), np.cos(az)
    return np.array([[cy * cz, sy * sx * cz - cx * sz, sy * cx * cz + sx * sz],
                     [cy * sz, sy * sx * sz + cx * cz, sy * cx * sz - sx * cz],
                     [-sy, cy * sx, cy * cx]])
