"""Oracle (test infrastructure): post-network pose math.

Restates
  * quaternion_matrix                   DenseFusion/lib/transformations.py:1254-1278
  * quaternion_from_matrix(isprecise)   DenseFusion/lib/transformations.py:1320-1341, :1361-1363
  * my_estimator_prediction             DenseFusion/tools/utils.py:7-18
  * my_refined_prediction               DenseFusion/tools/utils.py:20-40
  * get_new_points                      DenseFusion/tools/utils.py:43-86
Pinned by the doctest vectors of transformations.py and by golden vectors from
the imported reference (tests/golden/pose_math.npz).
"""
import math
import numpy as np

_EPS = np.finfo(float).eps * 4.0          # transformations.py:1669


def quaternion_matrix(q):
    """[w,x,y,z] -> 4x4 fp64 homogeneous rotation (transformations.py:1266-1278)."""
    q = np.array(q, dtype=np.float64)
    n = float(np.dot(q, q))
    if n < _EPS:
        return np.identity(4)
    q = q * math.sqrt(2.0 / n)
    o = np.outer(q, q)
    return np.array([
        [1.0 - o[2, 2] - o[3, 3], o[1, 2] - o[3, 0], o[1, 3] + o[2, 0], 0.0],
        [o[1, 2] + o[3, 0], 1.0 - o[1, 1] - o[3, 3], o[2, 3] - o[1, 0], 0.0],
        [o[1, 3] - o[2, 0], o[2, 3] + o[1, 0], 1.0 - o[1, 1] - o[2, 2], 0.0],
        [0.0, 0.0, 0.0, 1.0]])


def quaternion_from_matrix_precise(M):
    """The isprecise=True branch (transformations.py:1321-1341) + sign rule (:1361)."""
    M = np.asarray(M, dtype=np.float64)[:4, :4]
    q = np.empty(4)
    t = M[0, 0] + M[1, 1] + M[2, 2] + M[3, 3]
    if t > M[3, 3]:
        q[:] = (t, M[2, 1] - M[1, 2], M[0, 2] - M[2, 0], M[1, 0] - M[0, 1])
    else:
        i, j, k = 0, 1, 2
        if M[1, 1] > M[0, 0]:
            i, j, k = 1, 2, 0
        if M[2, 2] > M[i, i]:
            i, j, k = 2, 0, 1
        t = M[i, i] - (M[j, j] + M[k, k]) + M[3, 3]
        v = np.empty(4)
        v[i] = t
        v[j] = M[i, j] + M[j, i]
        v[k] = M[k, i] + M[i, k]
        v[3] = M[k, j] - M[j, k]
        q[:] = v[[3, 0, 1, 2]]
    q *= 0.5 / math.sqrt(t * M[3, 3])
    if q[0] < 0.0:
        q = -q
    return q


def estimator_prediction(pred_r, pred_t, pred_c, cloud):
    """tools/utils.py:7-18 on numpy fp32 inputs r[N,4], t[N,3], c[N], cloud[N,3].
    Returns (which_max, my_r fp32[4], my_t fp32[3])."""
    pred_r = np.asarray(pred_r, np.float32)
    nrm = np.sqrt((pred_r * pred_r).sum(axis=1, dtype=np.float32)).astype(np.float32)
    i = int(np.argmax(np.asarray(pred_c).reshape(-1)))
    my_r = (pred_r[i] / nrm[i]).astype(np.float32)
    my_t = (np.asarray(cloud, np.float32)[i] + np.asarray(pred_t, np.float32)[i]).astype(np.float32)
    return i, my_r, my_t


def refined_prediction(pred_r2, pred_t2, my_r, my_t):
    """tools/utils.py:20-40: fp64 compose M(my_r,my_t) @ M(r2/|r2|, t2).
    pred_r2 fp32[4] (un-normalised; normalised in fp32 as torch does), pred_t2 fp32[3]."""
    M1 = quaternion_matrix(my_r)
    M1[0:3, 3] = my_t
    r2 = np.asarray(pred_r2, np.float32).reshape(4)
    n2 = np.sqrt((r2 * r2).sum(dtype=np.float32)).astype(np.float32)
    r2 = (r2 / n2).astype(np.float32)
    M2 = quaternion_matrix(r2)
    M2[0:3, 3] = np.asarray(pred_t2, np.float32).reshape(3)
    Mf = np.dot(M1, M2)
    Rf = Mf.copy()
    Rf[0:3, 3] = 0
    q = quaternion_from_matrix_precise(Rf)
    t = np.array([Mf[0, 3], Mf[1, 3], Mf[2, 3]])
    return q, t


def quat_to_rowmajor_base(q):
    """tools/utils.py:46-67: 3x3 'base' from a normalised fp32 quaternion (fp32)."""
    q = np.asarray(q, np.float32)
    w, x, y, z = q
    two = np.float32(2.0)
    one = np.float32(1.0)
    return np.array([
        [one - two * (y * y + z * z), two * x * y - two * w * z, two * w * y + two * x * z],
        [two * x * y + two * z * w, one - two * (x * x + z * z), -two * w * x + two * y * z],
        [-two * w * y + two * x * z, two * w * x + two * y * z, one - two * (x * x + y * y)],
    ], dtype=np.float32)


def new_points(pred_r, pred_t, pred_c, points):
    """tools/utils.py:43-86 -> fp32 [N,3] = (points - t) @ base(argmax c)."""
    i, r, t = estimator_prediction(pred_r, pred_t, pred_c, points)
    base = quat_to_rowmajor_base(r)
    return ((np.asarray(points, np.float32) - t) @ base).astype(np.float32)


def rotation_angle_between(q1, q2):
    """Geodesic angle (rad) between two unit quaternions (sign-agnostic)."""
    q1 = np.asarray(q1, np.float64); q2 = np.asarray(q2, np.float64)
    d = abs(float(np.dot(q1 / np.linalg.norm(q1), q2 / np.linalg.norm(q2))))
    return 2.0 * math.acos(min(1.0, d))
