"""Generate tests/golden/kabsch_ref.npz by running the REFERENCE's own rigid SVD fit
`affine_matrix_from_points(v0, v1, shear=False, scale=False, usesvd=True)`
(DenseFusion/lib/transformations.py:889-995, imported unmodified from /root/reference) -- the Kabsch / Umeyama step that
open3d's TransformationEstimationPointToPoint performs inside registration_icp (open3d_utils.py:98-104).

Run in the dev container only:  python -m oracle.gen_golden_kabsch        TEST INFRASTRUCTURE (see oracle/__init__.py).

Cases (millimetre scale, as the label path): well-conditioned clouds of 3 .. 2000 points, noisy correspondences, a
reflection case (the optimal orthogonal matrix has det < 0 and the last singular direction must be flipped,
transformations.py:962-965), planar (rank-2) clouds with and without noise, and near-identity motions as ICP sees them.
Every case also stores whether nearest-neighbour correspondences within 10 mm are the identity pairing, so that the GPU test
can drive ONE iteration of the ICP kernel with it and compare the kernel's Kabsch update with the reference matrix.
"""
import os
import sys

import numpy as np

REF = os.environ.get('APE_REFERENCE', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')


def _rot(rng, max_angle):
    ax = rng.standard_normal(3); ax /= np.linalg.norm(ax)
    a = rng.uniform(-max_angle, max_angle)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    return np.identity(3) + np.sin(a) * K + (1 - np.cos(a)) * K @ K


def _lattice(rng, n, spacing=20.0, jitter=3.0, planar=False):
    """n points on a jittered lattice: minimum separation spacing - 2*jitter, so small motions keep NN == identity."""
    side = int(np.ceil(n ** (0.5 if planar else 1.0 / 3.0)))
    g = np.stack(np.meshgrid(*([np.arange(side)] * (2 if planar else 3)), indexing='ij'), -1).reshape(-1, 2 if planar else 3)
    g = g[rng.permutation(len(g))[:n]].astype(np.float64) * spacing
    if planar:
        g = np.concatenate([g, np.zeros((n, 1))], 1)
    g += rng.uniform(-jitter, jitter, size=g.shape) * ([1, 1, 0] if planar else [1, 1, 1])
    return g - g.mean(axis=0)


def main():
    sys.path.insert(0, REF)
    import DenseFusion.lib.transformations as tf
    rng = np.random.RandomState(909)
    cases = []

    def add(name, v0, v1):
        M = tf.affine_matrix_from_points(v0.T, v1.T, shear=False, scale=False, usesvd=True)
        d = np.linalg.norm(v0[:, None, :] - v1[None, :, :], axis=2)
        ident = bool((d.argmin(axis=1) == np.arange(len(v0))).all() and (d[np.arange(len(v0)), np.arange(len(v0))] < 10.0).all())
        cases.append((name, v0, v1, M, ident))

    for n in (3, 4, 20, 200, 2000):                              # exact rigid motions, ICP-sized steps
        P = _lattice(rng, n)
        R = _rot(rng, np.radians(2.0)); t = rng.uniform(-1.5, 1.5, size=3)
        add('rigid_n%d' % n, P, P @ R.T + t)
    for n in (50, 500):                                          # noisy correspondences
        P = _lattice(rng, n)
        R = _rot(rng, np.radians(3.0)); t = rng.uniform(-1, 1, size=3)
        add('noisy_n%d' % n, P, P @ R.T + t + rng.standard_normal(P.shape) * 0.3)
    P = _lattice(rng, 300)                                       # large motion (no identity NN; oracle-only case)
    add('large_motion', P, P @ _rot(rng, np.radians(120.0)).T + rng.uniform(-200, 200, size=3))
    P = _lattice(rng, 64, planar=True)                           # planar, exact
    add('planar_exact', P, P @ _rot(rng, np.radians(2.0)).T + rng.uniform(-1, 1, size=3))
    P = _lattice(rng, 64, planar=True)                           # reflection: thin slab whose z is mirrored
    Pz = P.copy(); Pz[:, 2] = rng.standard_normal(64) * 0.05
    Q = Pz.copy(); Q[:, 2] *= -1.0
    add('reflection', Pz, Q @ _rot(rng, np.radians(1.0)).T + rng.uniform(-0.5, 0.5, size=3))
    P = _lattice(rng, 100, planar=True)                          # planar + noise on both sides
    add('planar_noisy', P + rng.standard_normal(P.shape) * [0.2, 0.2, 0.0], P @ _rot(rng, np.radians(1.5)).T + rng.standard_normal(P.shape) * 0.2)
    out = dict(names=np.array([c[0] for c in cases]), identity_nn=np.array([c[4] for c in cases]))
    for i, (_, v0, v1, M, _) in enumerate(cases):
        out['v0_%d' % i] = v0; out['v1_%d' % i] = v1; out['M_%d' % i] = M
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, 'kabsch_ref.npz'), **out)
    for c in cases:
        print('%-14s n=%5d det=%+.6f identity_nn=%s' % (c[0], len(c[1]), np.linalg.det(c[3][:3, :3]), c[4]))


if __name__ == '__main__':
    main()
