"""CPU: pin the oracle against (a) the reference's own doctest vectors, (b) golden vectors
produced by the imported reference modules (oracle/gen_golden.py), (c) the reference's own
knn_cpu.cpp binary when oracle/_ref is present."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import clib, densefusion as odf, icp as oicp, pose_math as pm, synth


def _g(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


# ---- (a) known-answer vectors from DenseFusion/lib/transformations.py doctests
def test_quaternion_matrix_doctests():
    # transformations.py:1257-1265
    M = pm.quaternion_matrix([0.99810947, 0.06146124, 0, 0])
    c, s = math.cos(0.123), math.sin(0.123)
    assert np.allclose(M[:3, :3], [[1, 0, 0], [0, c, -s], [0, s, c]])
    assert np.allclose(pm.quaternion_matrix([1, 0, 0, 0]), np.identity(4))
    assert np.allclose(pm.quaternion_matrix([0, 1, 0, 0]), np.diag([1, -1, -1, 1]))


def test_quaternion_from_matrix_doctests():
    # transformations.py:1287-1295
    assert np.allclose(pm.quaternion_from_matrix_precise(np.identity(4)), [1, 0, 0, 0])
    ax = np.array([1.0, 2.0, 3.0]); ax /= np.linalg.norm(ax)
    a = 0.123
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.identity(4); R[:3, :3] = np.identity(3) + math.sin(a) * K + (1 - math.cos(a)) * K @ K
    assert np.allclose(pm.quaternion_from_matrix_precise(R), [0.9981095, 0.0164262, 0.0328524, 0.0492786])


def test_kabsch_matches_reference_fixture(golden_dir):
    """The ICP update step (Eigen umeyama without scaling inside open3d's point-to-point estimation) against the
    reference's own rigid SVD fit, affine_matrix_from_points(v0, v1, shear=False, scale=False)
    (DenseFusion/lib/transformations.py:889-995), run by oracle/gen_golden_kabsch.py: exact and noisy correspondences,
    the reflection branch (:962-965), rank-2 (planar, 3-point) inputs, a large motion."""
    g = _g(golden_dir, 'kabsch_ref.npz')
    names = list(g['names'])
    assert {'reflection', 'planar_exact', 'rigid_n3', 'large_motion'} <= set(names)
    for i, name in enumerate(names):
        v0, v1, M = g['v0_%d' % i], g['v1_%d' % i], g['M_%d' % i]
        T = oicp.kabsch_umeyama(v0, v1)
        assert np.abs(T - M).max() < 1e-12 * max(1.0, np.abs(M).max()), (name, np.abs(T - M).max())
        assert abs(np.linalg.det(T[:3, :3]) - 1.0) < 1e-12, name


def test_kabsch_recovers_known_motion():
    rng = np.random.RandomState(0)
    P = rng.rand(20, 3) - 0.5
    R = synth.random_rotation(rng, 1.0); t = rng.rand(3)
    T = oicp.kabsch_umeyama(P, P @ R.T + t)
    assert np.allclose(T[:3, :3], R, atol=1e-12) and np.allclose(T[:3, 3], t, atol=1e-12)


def test_compiled_reference_modules_match_golden(golden_dir):
    """oracle/_ref/DenseFusion/*.so (the reference's own Python modules compiled by Cython, used as bench.py's CPU arm)
    reproduce the fixture that the imported reference sources produced, and agree with the oracle restatement."""
    from oracle import ref_modules
    if not ref_modules.available():
        pytest.skip('oracle/_ref/DenseFusion not built (needs /root/reference at build time)')
    mods = ref_modules.load()
    network = mods[0]
    g = _g(golden_dir, 'densefusion_case0.npz')
    seed, npts, nobj = int(g['seed']), int(g['npts']), int(g['nobj'])
    hw = tuple(int(v) for v in g['hw'])
    sd_e = synth.to_torch(synth.posenet_state_dict(seed, nobj)); sd_r = synth.to_torch(synth.refiner_state_dict(seed + 1000, nobj))
    est = network.PoseNet(npts, nobj); est.cnn = torch.nn.Identity(); est.eval(); est.load_state_dict(sd_e, strict=False)
    refn = network.PoseRefineNet(npts, nobj); refn.eval(); refn.load_state_dict(sd_r)
    t = [torch.from_numpy(a) for a in synth.posenet_inputs(seed, npts, hw, nobj)]
    torch.set_num_threads(4)
    with torch.no_grad():
        r, tt, c, emb = est(*t)
        assert np.array_equal(r.numpy(), g['r']) and np.array_equal(tt.numpy(), g['t']) and np.array_equal(c.numpy(), g['c'])
        q, tr = ref_modules.canonical_prediction(mods, est, refn, t[0], t[1], t[2], t[3], npts, 2)
        res = odf.canonical_prediction(sd_e, sd_r, t[0], t[1], t[2], t[3], nobj, iterations=2)
    assert pm.rotation_angle_between(q, res['q']) < 1e-6 and np.abs(tr - res['t']).max() < 1e-6


# ---- (b) golden vectors from the imported reference
def test_pose_math_golden(golden_dir):
    g = _g(golden_dir, 'pose_math.npz')
    for q, M in zip(g['quats'], g['quat_mats']):
        assert np.array_equal(pm.quaternion_matrix(q), M)
    for R, q in zip(g['rots'], g['rot_quats']):
        assert np.allclose(pm.quaternion_from_matrix_precise(R), q, rtol=0, atol=1e-15)
    for i in range(len(g['r2'])):
        q, t = pm.refined_prediction(g['r2'][i], g['t2'][i], g['my_r'][i], g['my_t'][i])
        assert np.allclose(q, g['refined_q'][i], atol=1e-12) and np.allclose(t, g['refined_t'][i], atol=1e-12)
    i, r, t = pm.estimator_prediction(g['pr'][0], g['pt'][0], g['pc'][0, :, 0], g['pts'][0])
    assert np.allclose(r, g['est_r'], atol=1e-7) and np.allclose(t, g['est_t'], atol=1e-7)
    npn = pm.new_points(g['pr'][0], g['pt'][0], g['pc'][0, :, 0], g['pts'][0])
    assert np.allclose(npn, g['new_points'][0], atol=2e-6)


@pytest.mark.parametrize('case', [0, 1, 2])
def test_densefusion_golden(golden_dir, case):
    g = _g(golden_dir, 'densefusion_case%d.npz' % case)
    seed, npts, nobj = int(g['seed']), int(g['npts']), int(g['nobj'])
    hw = tuple(int(v) for v in g['hw'])
    sd_e = synth.to_torch(synth.posenet_state_dict(seed, nobj))
    sd_r = synth.to_torch(synth.refiner_state_dict(seed + 1000, nobj))
    out_img, cloud, choose, idx = (torch.from_numpy(a) for a in synth.posenet_inputs(seed, npts, hw, nobj))
    torch.set_num_threads(4)
    with torch.no_grad():
        r, t, c, emb = odf.posenet_geometry(sd_e, out_img, cloud, choose, idx, nobj)
        assert np.allclose(r.numpy(), g['r'], atol=1e-5, rtol=1e-5)
        assert np.allclose(t.numpy(), g['t'], atol=1e-5, rtol=1e-5)
        assert np.allclose(c.numpy(), g['c'], atol=1e-6)
        assert np.allclose(emb.numpy().sum(axis=1), g['emb_sum'], atol=1e-4)
        res = odf.live_prediction(sd_e, sd_r, out_img, cloud, choose, idx, nobj)
    assert np.allclose(res['my_r'], g['my_r'], atol=1e-6) and np.allclose(res['my_t'], g['my_t'], atol=1e-6)
    assert np.allclose(res['r2'], g['r2'][0], atol=1e-5) and np.allclose(res['t2'], g['t2'][0], atol=1e-5)
    assert pm.rotation_angle_between(res['q'], g['final_q']) < 1e-5
    assert np.allclose(res['t'], g['final_t'], atol=1e-6)


def test_losses_golden(golden_dir):
    g = _g(golden_dir, 'losses.npz')
    T = lambda k: torch.from_numpy(g[k])
    for tag, sym in (('sym', [0]), ('nosym', [])):
        dis, npn, ntg, pred = odf.loss_refine(T('pr1'), T('pt1'), T('target'), T('model'), torch.LongTensor([[0]]), T('points'), sym)
        assert np.allclose(dis.numpy(), g['lr_dis_' + tag], atol=1e-7)
        assert np.allclose(npn.numpy(), g['lr_newp_' + tag], atol=1e-6)
        assert np.allclose(ntg.numpy(), g['lr_newt_' + tag], atol=1e-6)
        assert np.allclose(pred.numpy(), g['lr_pred_' + tag], atol=1e-6)
        # C restatement of the metric (what the CUDA kernel is checked against)
        d_c = clib.add_metric(g['pr1'][0], g['pt1'][0], g['model'][0], g['target'][0], bool(sym))
        assert abs(d_c - float(g['lr_dis_' + tag][0])) < 1e-6
    for tag, sym, refine in (('sym', [0], False), ('nosym', [], False), ('symrefine', [0], True)):
        lo, dis, npn, ntg, _ = odf.loss_estimator(T('pr_n'), T('pt_n'), T('pc_n'), T('target'), T('model'),
                                                  torch.LongTensor([[0]]), T('points'), 0.015, refine, sym)
        assert np.allclose(lo.numpy(), g['l_loss_' + tag], atol=1e-6)
        assert np.allclose(dis.numpy(), g['l_dis_' + tag], atol=1e-6)
        assert np.allclose(npn.numpy(), g['l_newp_' + tag], atol=1e-6)
        assert np.allclose(ntg.numpy(), g['l_newt_' + tag], atol=1e-6)


def test_knn_golden(golden_dir):
    g = _g(golden_dir, 'knn.npz')
    assert np.array_equal(clib.knn(g['ref'], g['qry'], 1, 0), g['idx_k1'])
    assert np.array_equal(clib.knn(g['ref'], g['qry'], 4, 0), g['idx_k4'])
    assert np.array_equal(clib.knn(g['refd'], g['qryd'], 2, 0), g['idxd_k2'])
    # numpy restatement used inside the loss oracle
    j = odf.knn_top1_np(g['ref'][0].T, g['qry'][0].T)
    assert np.array_equal(j + 1, g['idx_k1'][0, 0])
    # exact tie: query 3 equals ref 5 and ref 17 -> lowest index (1-based 6)
    assert g['idx_k1'][0, 0, 3] == 6


# ---- (c) live against the reference's own compiled knn_cpu.cpp (dev container / travels in oracle/_ref)
def test_knn_against_reference_binary():
    rng = np.random.RandomState(5)
    ref = rng.rand(1, 3, 150).astype(np.float32); qry = rng.rand(1, 3, 40).astype(np.float32)
    out = clib.ref_knn_cpu(ref, qry, 3)
    if out is None:
        pytest.skip('oracle/_ref not built (reference tree absent)')
    assert np.array_equal(out, clib.knn(ref, qry, 3, 0))


# ---- ICP restatement self-consistency (parity unpinned at the open3d boundary)
def test_icp_oracle_recovers_known_motion():
    rng = np.random.RandomState(7)
    tgt = synth.ellipsoid_cloud(rng, 1500)
    R = synth.random_rotation(rng, math.radians(6)); t = np.array([1.5, -2.0, 1.0])
    src = (tgt[rng.choice(1500, 900, replace=False)] - t) @ R     # R^T (q - t)
    T, info = oicp.registration_icp_p2p(src, tgt, 10.0, return_info=True)
    assert info['fitness'] > 0.99 and info['inlier_rmse'] < 2.0
    T2 = oicp.registration_icp_p2p(src, tgt, 10.0, use_kdtree=True)
    assert np.allclose(T, T2, atol=1e-9)
    # criteria semantics: stops well before max_iteration
    assert info['iterations'] < 100


def test_voxel_down_sample_oracle():
    p = np.array([[0.1, 0.1, 0.1], [0.2, 0.3, 0.1], [4.2, 4.1, 4.3], [4.4, 4.5, 4.6], [9.9, 0.0, 0.0]])
    v = oicp.voxel_down_sample(p, 2.0)
    assert v.shape == (3, 3)
    assert np.allclose(sorted(v[:, 0]), sorted([0.15, 4.3, 9.9]))


# ---- (d) geometry oracle against vectors produced by the reference's OWN dataset / get_surface code
#      (oracle/gen_golden_geometry.py: DenseFusion/datasets/myDatasetAugmented/dataset.py, pc_reconstruction/open3d_utils.py)
def test_geometry_bbox_golden(golden_dir):
    from oracle import geometry as og
    g = _g(golden_dir, 'geometry_ref.npz')
    for (r0, r1, c0, c1), want in zip(g['rects'], g['rect_bbox']):
        m = np.zeros((480, 640), bool); m[r0:r1, c0:c1] = True
        assert tuple(int(v) for v in og.get_bbox(m)) == tuple(int(v) for v in want)
    for i in range(len(g['bbox'])):
        assert tuple(int(v) for v in og.get_bbox(g['label'][i] == 255)) == tuple(int(v) for v in g['bbox'][i])


def test_geometry_choose_and_backprojection_golden(golden_dir):
    """dataset.py:236-275 run by the reference itself (np.random.seed(s) before each sample): the oracle must reproduce
    `choose` exactly (shuffle branch for frames 0/1, 'wrap' branch for frame 2) and the fp32 cloud bit for bit."""
    from oracle import geometry as og
    g = _g(golden_dir, 'geometry_ref.npz')
    ppx, ppy, fx, fy = (float(v) for v in g['intr'])
    N = int(g['num_pt'])
    branches = set()
    for i, s in enumerate(g['seeds']):
        depth, mask = g['depth'][i], g['label'][i] == 255
        bbox = og.get_bbox(mask)
        cand = og.choose_candidates(mask, depth, bbox)
        keep = og.make_keep(len(cand), N, np.random.RandomState(int(s))) if len(cand) > N else None
        branches.add(keep is None)
        choose = og.choose_fixed(cand, N, keep)
        assert np.array_equal(choose, g['choose'][i])
        cloud = og.backproject_choose(depth, bbox, choose, ppx, ppy, fx, fy, float(g['depth_scale']))
        assert cloud.dtype == np.float32 and np.array_equal(cloud.view(np.uint32), g['cloud'][i].view(np.uint32))
    assert branches == {True, False}


def test_geometry_surface_golden(golden_dir):
    """open3d_utils.py:171-192 run by the reference itself: same points, same (row-major) order, bit for bit in fp64 for
    the literal restatement; the vectorised form (what the GPU tests use at full size) within 1e-9 mm."""
    from oracle import geometry as og
    g = _g(golden_dir, 'geometry_ref.npz')
    intr = dict(zip(('ppx', 'ppy', 'fx', 'fy'), (float(v) for v in g['intr'])))
    off = np.concatenate([[0], np.cumsum(g['surf_n'])])
    for i in range(len(g['surf_n'])):
        want = g['surf_pts'][off[i]:off[i + 1]]
        depth = g['depth'][i].astype(np.float64)
        vec, pix = og.surface_backproject(g['label'][i], depth, intr, g['robot2cam'])
        assert vec.shape == want.shape and np.abs(vec - want).max() < 1e-9
        if i == 2:      # small frame: the literal per-pixel loop is cheap
            lit, pix_l = og.surface_backproject_literal(g['label'][i], depth, intr, g['robot2cam'])
            assert np.array_equal(lit, want) and np.array_equal(pix, pix_l)
