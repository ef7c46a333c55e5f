"""Oracle (test infrastructure): mask -> bbox -> choose -> back-projection.

numpy restatements of
  * get_bbox                  DenseFusion/datasets/myDatasetAugmented/dataset.py:338-380
  * mask / choose             pipeline/utils.py:524-539 (= dataset.py:239-257)
  * fp32 back-projection      pipeline/utils.py:542-553
  * get_surface projection    pc_reconstruction/open3d_utils.py:171-192
Pinned by tests/golden/geometry_ref.npz = outputs of the reference's own dataset.py / open3d_utils.py run on seeded
frames (oracle/gen_golden_geometry.py; tests/test_oracle_pinned.py).
"""
import numpy as np

IMG_H = 480   # dataset.py:339 ("img_width")
IMG_W = 640   # dataset.py:340 ("img_length")
BORDER = 40   # dataset.py:338: border_list = [-1, 40, 80, ..., 680]


def _round_up_border(extent):
    """dataset.py:350-358: an extent strictly between two border_list entries is
    raised to the upper entry; exact multiples of 40 are kept."""
    edges = [-1] + [BORDER * i for i in range(1, 18)]
    for lo, hi in zip(edges[:-1], edges[1:]):
        if lo < extent < hi:
            return hi
    return extent


def get_bbox(mask):
    """dataset.py:342-380.  mask: bool/uint8 [480,640] -> (rmin, rmax, cmin, cmax)."""
    mask = np.asarray(mask) != 0
    r_any = np.flatnonzero(mask.any(axis=1))
    c_any = np.flatnonzero(mask.any(axis=0))
    rmin, rmax = int(r_any[0]), int(r_any[-1]) + 1
    cmin, cmax = int(c_any[0]), int(c_any[-1]) + 1
    r_b = _round_up_border(rmax - rmin)
    c_b = _round_up_border(cmax - cmin)
    cr, cc = int((rmin + rmax) / 2), int((cmin + cmax) / 2)
    rmin, rmax = cr - int(r_b / 2), cr + int(r_b / 2)
    cmin, cmax = cc - int(c_b / 2), cc + int(c_b / 2)
    if rmin < 0:
        rmax, rmin = rmax - rmin, 0
    if cmin < 0:
        cmax, cmin = cmax - cmin, 0
    if rmax > IMG_H:
        rmin, rmax = rmin - (rmax - IMG_H), IMG_H
    if cmax > IMG_W:
        cmin, cmax = cmin - (cmax - IMG_W), IMG_W
    return rmin, rmax, cmin, cmax


def choose_candidates(label_mask, depth, bbox):
    """pipeline/utils.py:524-529: row-major flat indices (inside the bbox crop) of
    pixels with label==255 and depth!=0."""
    rmin, rmax, cmin, cmax = bbox
    m = (np.asarray(label_mask) != 0) & (np.asarray(depth) != 0)
    return np.flatnonzero(m[rmin:rmax, cmin:cmax].ravel())


def choose_fixed(candidates, num_points, keep=None):
    """pipeline/utils.py:530-539.  ``keep`` is the shuffled 0/1 vector the reference
    draws with np.random.shuffle (the "fixed sampling indices" of the contract);
    when there are at most num_points candidates the reference wraps (np.pad)."""
    n = len(candidates)
    if n == 0:
        return None                       # reference: `continue` (:530-531)
    if n > num_points:
        assert keep is not None and keep.shape[0] == n and int(keep.sum()) == num_points
        return candidates[np.flatnonzero(keep)]
    return np.pad(candidates, (0, num_points - n), 'wrap')


def make_keep(n, num_points, rng):
    """The reference's c_mask (:533-536) with an explicit RandomState."""
    k = np.zeros(n, dtype=int)
    k[:num_points] = 1
    rng.shuffle(k)
    return k


def backproject_choose(depth, bbox, choose, ppx, ppy, fx, fy, depth_scale):
    """pipeline/utils.py:542-553.  fp32, left-to-right, python scalars rounded to
    fp32 (numpy weak-scalar promotion).  Returns float32 [N,3] = (x, y, z).
    NOTE xmap holds the ROW index and ymap the COLUMN index (:518-519)."""
    rmin, rmax, cmin, cmax = bbox
    width = cmax - cmin
    choose = np.asarray(choose).astype(np.int64)
    rows = (rmin + choose // width).astype(np.float32)[:, None]
    cols = (cmin + choose % width).astype(np.float32)[:, None]
    d = np.asarray(depth)[rmin:rmax, cmin:cmax].ravel()[choose][:, None].astype(np.float32)
    f32 = np.float32
    pt2 = d * f32(depth_scale)
    pt0 = (cols - f32(ppx)) * pt2 / f32(fx)
    pt1 = (rows - f32(ppy)) * pt2 / f32(fy)
    return np.concatenate((pt0, pt1, pt2), axis=1).astype(np.float32)


def surface_backproject_literal(label, depth_frame, intr, robot2cam):
    """open3d_utils.py:172-192 as written: per-pixel loop, 4x4 np.dot per pixel.
    Returns (points float64 [n,3], flat pixel index int64 [n]).  Slow (~10 us/px)."""
    pts, pix = [], []
    rr, cc = np.where(np.asarray(label) != 0)
    for r, c in zip(rr, cc):
        z = depth_frame[r, c]
        if z != 0:
            x = (c - intr['ppx']) * z / intr['fx']
            y = (r - intr['ppy']) * z / intr['fy']
            cam2obj = np.identity(4)
            cam2obj[0:3, 3] = [x, y, z]
            pts.append(np.dot(robot2cam, cam2obj)[:3, 3])
            pix.append(r * label.shape[1] + c)
    if not pts:
        return np.zeros((0, 3)), np.zeros((0,), dtype=np.int64)
    return np.array(pts, dtype=np.float64), np.array(pix, dtype=np.int64)


def surface_backproject(label, depth_frame, intr, robot2cam):
    """Vectorised form of the same arithmetic (fp64; same per-element op order for
    x, y; the rigid transform is R*p + t evaluated row by row)."""
    label = np.asarray(label)
    depth_frame = np.asarray(depth_frame, dtype=np.float64)
    sel = (label != 0) & (depth_frame != 0)
    rr, cc = np.nonzero(sel)                      # row-major order, as np.where (:174)
    z = depth_frame[rr, cc]
    x = (cc - intr['ppx']) * z / intr['fx']
    y = (rr - intr['ppy']) * z / intr['fy']
    T = np.asarray(robot2cam, dtype=np.float64)
    out = np.empty((len(z), 3), dtype=np.float64)
    for i in range(3):
        out[:, i] = T[i, 0] * x + T[i, 1] * y + T[i, 2] * z + T[i, 3]
    return out, (rr * label.shape[1] + cc).astype(np.int64)


# ----------------------------------------------------------------------------- device-side sampling
def mix32(x):
    """murmur3 finaliser on uint32 arrays (csrc/choose.cu: mix32)."""
    x = np.asarray(x, np.uint64) & 0xffffffff
    x ^= x >> 16; x = (x * 0x85ebca6b) & 0xffffffff
    x ^= x >> 13; x = (x * 0xc2b2ae35) & 0xffffffff
    x ^= x >> 16
    return x.astype(np.uint32)


def choose_hashed(candidates, num_points, seed):
    """The device's replacement for the np.random.shuffle subset of pipeline/utils.py:532-537: candidate c gets the key
    mix32(seed ^ c * 0x9E3779B9); the num_points smallest keys are kept (ties: lower index first), in ascending index
    order.  At most num_points candidates -> 'wrap' padding (:539); none -> None (:530-531)."""
    n = len(candidates)
    if n == 0:
        return None
    if n <= num_points:
        return np.pad(candidates, (0, num_points - n), 'wrap')
    c = np.asarray(candidates, np.uint64)
    keys = mix32((np.uint64(seed & 0xffffffff) ^ ((c * np.uint64(0x9E3779B9)) & np.uint64(0xffffffff))))
    order = np.lexsort((np.asarray(candidates), keys))            # by key, then by index
    return np.sort(np.asarray(candidates)[order[:num_points]])
