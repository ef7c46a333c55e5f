"""GPU parity: confidence arg-max / pose select / new_points and the fp64 pose composition."""
import os

import numpy as np
import pytest
import torch

from oracle import pose_math as pm

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_pose_select_golden(golden_dir):
    from autoposeestimation_b200 import ops
    g = np.load(os.path.join(golden_dir, 'pose_math.npz'))
    out = ops.pose_select(_dev(g['pr']), _dev(g['pt']), _dev(g['pc']), _dev(g['pts']))
    i, r, t = pm.estimator_prediction(g['pr'][0], g['pt'][0], g['pc'][0, :, 0], g['pts'][0])
    assert int(out['which_max'].cpu()[0]) == i
    assert np.allclose(out['my_r'].cpu().numpy()[0], g['est_r'], atol=1e-7)          # reference my_estimator_prediction
    assert np.allclose(out['my_t'].cpu().numpy()[0], g['est_t'], atol=1e-7)
    assert np.allclose(out['new_points'].cpu().numpy(), g['new_points'], atol=2e-6)    # reference get_new_points
    assert np.allclose(out['pose'].cpu().numpy()[0], np.concatenate([g['est_r'], g['est_t']]), atol=1e-7)


def test_pose_select_batch_ties_lowest_index():
    from autoposeestimation_b200 import ops
    rng = np.random.RandomState(4)
    B, N = 7, 1000
    pr = rng.standard_normal((B, N, 4)).astype(np.float32); pt = rng.standard_normal((B, N, 3)).astype(np.float32) * 0.1
    pc = rng.uniform(0.1, 0.9, size=(B, N)).astype(np.float32); pts = rng.standard_normal((B, N, 3)).astype(np.float32)
    pc[0, 700] = pc[0, 123] = 0.95                      # exact tie -> lowest index
    pc[1, 0] = 0.99; pc[2, N - 1] = 0.99
    out = ops.pose_select(_dev(pr), _dev(pt), _dev(pc), _dev(pts))
    wm = out['which_max'].cpu().numpy()
    assert wm[0] == 123 and wm[1] == 0 and wm[2] == N - 1
    for b in range(B):
        i, r, t = pm.estimator_prediction(pr[b], pt[b], pc[b], pts[b])
        assert wm[b] == i
        assert np.allclose(out['my_r'].cpu().numpy()[b], r, atol=1e-7) and np.allclose(out['my_t'].cpu().numpy()[b], t, atol=1e-7)
        assert np.allclose(out['new_points'].cpu().numpy()[b], pm.new_points(pr[b], pt[b], pc[b], pts[b]), atol=3e-6)


def test_pose_compose_golden(golden_dir):
    """my_refined_prediction outputs of the reference (tests/golden/pose_math.npz), all quaternion branches."""
    from autoposeestimation_b200 import ops
    g = np.load(os.path.join(golden_dir, 'pose_math.npz'))
    pose_in = np.concatenate([g['my_r'].astype(np.float64), g['my_t'].astype(np.float64)], axis=1)
    out = ops.pose_compose(_dev(pose_in), _dev(g['r2']), _dev(g['t2'])).cpu().numpy()
    assert np.allclose(out[:, :4], g['refined_q'], atol=1e-12)
    assert np.allclose(out[:, 4:], g['refined_t'], atol=1e-12)


def test_pose_compose_branches_and_next_points():
    from autoposeestimation_b200 import ops
    rng = np.random.RandomState(8)
    B, N = 40, 333
    my_r = rng.standard_normal((B, 4)); my_r /= np.linalg.norm(my_r, axis=1, keepdims=True)
    # near-180-degree rotations hit the non-trace branches of quaternion_from_matrix
    for b, ax in enumerate(([1, 0, 0], [0, 1, 0], [0, 0, 1])):
        a = np.pi - 1e-3 * (b + 1)
        my_r[b] = [np.cos(a / 2)] + list(np.sin(a / 2) * np.array(ax, float))
    my_t = rng.standard_normal((B, 3)) * 0.3
    r2 = rng.standard_normal((B, 4)).astype(np.float32); r2[:3] = [[1, 1e-4, 0, 0], [1, 0, 1e-4, 0], [1, 0, 0, 1e-4]]
    t2 = (rng.standard_normal((B, 3)) * 0.02).astype(np.float32)
    cloud = rng.standard_normal((B, N, 3)).astype(np.float32)
    pose_in = np.concatenate([my_r, my_t], axis=1)
    out, nxt = ops.pose_compose(_dev(pose_in), _dev(r2), _dev(t2), _dev(cloud))
    out = out.cpu().numpy(); nxt = nxt.cpu().numpy()
    for b in range(B):
        q, t = pm.refined_prediction(r2[b], t2[b], my_r[b], my_t[b])
        assert np.allclose(out[b, :4], q, atol=1e-12) and np.allclose(out[b, 4:], t, atol=1e-12)
        R = pm.quaternion_matrix(q)[:3, :3].astype(np.float32)                 # eval_linemod.py:92-97
        want = (cloud[b] - t.astype(np.float32)) @ R
        assert np.allclose(nxt[b], want, atol=3e-6)
