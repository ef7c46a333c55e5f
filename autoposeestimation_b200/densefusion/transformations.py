"""The two functions of DenseFusion/lib/transformations.py that the hot path uses (host side, fp64 numpy):
quaternion_matrix (:1254-1278) and quaternion_from_matrix (:1281-1363).  The device versions live in
csrc/pose_math.cu; these exist so callers that import them from the drop-in keep working."""
import math

import numpy

_EPS = numpy.finfo(float).eps * 4.0


def quaternion_matrix(quaternion):
    q = numpy.array(quaternion, dtype=numpy.float64)
    n = float(numpy.dot(q, q))
    if n < _EPS:
        return numpy.identity(4)
    w, x, y, z = q * math.sqrt(2.0 / n)
    return numpy.array([[1.0 - y * y - z * z, x * y - z * w, x * z + y * w, 0.0],
                        [x * y + z * w, 1.0 - x * x - z * z, y * z - x * w, 0.0],
                        [x * z - y * w, y * z + x * w, 1.0 - x * x - y * y, 0.0],
                        [0.0, 0.0, 0.0, 1.0]])


def quaternion_from_matrix(matrix, isprecise=False):
    M = numpy.asarray(matrix, dtype=numpy.float64)[:4, :4]
    if isprecise:
        t = M[0, 0] + M[1, 1] + M[2, 2] + M[3, 3]
        if t > M[3, 3]:
            q = numpy.array([t, M[2, 1] - M[1, 2], M[0, 2] - M[2, 0], M[1, 0] - M[0, 1]])
        else:
            i, j, k = 0, 1, 2
            if M[1, 1] > M[0, 0]:
                i, j, k = 1, 2, 0
            if M[2, 2] > M[i, i]:
                i, j, k = 2, 0, 1
            t = M[i, i] - (M[j, j] + M[k, k]) + M[3, 3]
            v = numpy.empty(4)
            v[i], v[j], v[k], v[3] = t, M[i, j] + M[j, i], M[k, i] + M[i, k], M[k, j] - M[j, k]
            q = v[[3, 0, 1, 2]]
        q = q * (0.5 / math.sqrt(t * M[3, 3]))
    else:
        m = M
        K = numpy.array([[m[0, 0] - m[1, 1] - m[2, 2], 0.0, 0.0, 0.0],
                         [m[0, 1] + m[1, 0], m[1, 1] - m[0, 0] - m[2, 2], 0.0, 0.0],
                         [m[0, 2] + m[2, 0], m[1, 2] + m[2, 1], m[2, 2] - m[0, 0] - m[1, 1], 0.0],
                         [m[2, 1] - m[1, 2], m[0, 2] - m[2, 0], m[1, 0] - m[0, 1], m[0, 0] + m[1, 1] + m[2, 2]]]) / 3.0
        w, V = numpy.linalg.eigh(K)
        q = V[[3, 0, 1, 2], numpy.argmax(w)]
    return -q if q[0] < 0.0 else q
