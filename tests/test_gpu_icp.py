"""GPU parity: persistent ICP kernel and voxel grid vs the open3d-0.9-semantics oracle (transforms within 1e-5)."""
import math

import numpy as np
import pytest
import torch

from oracle import geometry as og, icp as oicp, synth

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _ragged(clouds):
    off = np.zeros(len(clouds) + 1, np.int32)
    off[1:] = np.cumsum([len(c) for c in clouds])
    return np.concatenate(clouds, axis=0) if clouds else np.zeros((0, 3)), off


def _surface(seed):
    fr = synth.render_ellipsoid_frame(seed)
    pts, _ = og.surface_backproject(fr['label'], fr['depth'].astype(np.float64), fr['intr'], fr['robot2cam'])
    return oicp.voxel_down_sample(pts, 2.0), fr['model']


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_icp_c1_vs_oracle(seed):
    """BASELINE config 1: one 640x480 frame back-projected + masked, registered to a 2000-pt model cloud."""
    from autoposeestimation_b200 import ops
    src, tgt = _surface(seed)
    T_ref, info_ref = oicp.registration_icp_p2p(src, tgt, 10.0, return_info=True)
    so, to = np.array([0, len(src)], np.int32), np.array([0, len(tgt)], np.int32)
    T, info = ops.icp_p2p(_dev(src), _dev(so), _dev(tgt), _dev(to), 10.0)
    T = T.cpu().numpy()[0]; info = info.cpu().numpy()[0]
    assert int(info[2]) == info_ref['iterations'] and int(info[3]) == info_ref['n_corr']
    assert abs(info[0] - info_ref['fitness']) < 1e-12 and abs(info[1] - info_ref['inlier_rmse']) < 1e-9
    assert np.abs(T - T_ref).max() < 1e-5, np.abs(T - T_ref).max()                 # the parity gate
    assert np.allclose(T[3], [0, 0, 0, 1])


def test_icp_kabsch_update_matches_reference_fixture(golden_dir):
    """ONE iteration of the ICP kernel on correspondences that are the identity pairing = its Kabsch / Umeyama update,
    against the reference's own rigid SVD fit (transformations.py:889-995, tests/golden/kabsch_ref.npz): exact and noisy
    pairs, the reflection branch, rank-2 clouds."""
    import os
    from autoposeestimation_b200 import ops
    g = np.load(os.path.join(golden_dir, 'kabsch_ref.npz'))
    idx = [i for i, ok in enumerate(g['identity_nn']) if ok]
    assert len(idx) >= 9
    srcs, tgts = [g['v0_%d' % i] for i in idx], [g['v1_%d' % i] for i in idx]
    S, so = _ragged(srcs); Tg, to = _ragged(tgts)
    T, info = ops.icp_p2p(_dev(S), _dev(so), _dev(Tg), _dev(to), 10.0, max_iter=1)
    T = T.cpu().numpy(); info = info.cpu().numpy()
    for k, i in enumerate(idx):
        M = g['M_%d' % i]
        assert int(info[k, 2]) == 1
        assert np.abs(T[k] - M).max() < 1e-9, (str(g['names'][i]), np.abs(T[k] - M).max())
        assert abs(np.linalg.det(T[k][:3, :3]) - 1.0) < 1e-9


def test_icp_ragged_batch_init_and_edge_cases():
    from autoposeestimation_b200 import ops
    rng = np.random.RandomState(21)
    srcs, tgts, inits = [], [], []
    for r in range(6):
        tgt = synth.ellipsoid_cloud(rng, 300 + 450 * r, (40.0 + 5 * r, 30.0, 20.0))
        R = synth.random_rotation(rng, math.radians(8)); t = rng.uniform(-3, 3, size=3)
        src = (tgt[rng.choice(len(tgt), 200 + 100 * r, replace=False)] - t) @ R + rng.standard_normal((200 + 100 * r, 3)) * 0.2
        srcs.append(src); tgts.append(tgt)
        I = np.identity(4)
        if r % 2:
            I[:3, :3] = synth.random_rotation(rng, math.radians(2)); I[:3, 3] = rng.uniform(-1, 1, size=3)
        inits.append(I)
    srcs.append(srcs[0] + 500.0); tgts.append(tgts[0]); inits.append(np.identity(4))      # no correspondences at all
    srcs.append(srcs[1][:1]); tgts.append(tgts[1]); inits.append(np.identity(4))            # single source point
    S, so = _ragged(srcs); Tg, to = _ragged(tgts)
    T, info = ops.icp_p2p(_dev(S), _dev(so), _dev(Tg), _dev(to), 10.0, init=_dev(np.stack(inits)))
    T = T.cpu().numpy(); info = info.cpu().numpy()
    for r in range(len(srcs)):
        T_ref, ir = oicp.registration_icp_p2p(srcs[r], tgts[r], 10.0, init=inits[r], return_info=True)
        assert int(info[r, 2]) == ir['iterations'], (r, info[r], ir)
        assert np.abs(T[r] - T_ref).max() < 1e-5, (r, np.abs(T[r] - T_ref).max())
    assert info[6, 0] == 0.0 and np.allclose(T[6], np.identity(4))


def test_icp_criteria_and_max_iteration():
    from autoposeestimation_b200 import ops
    src, tgt = _surface(3)
    so, to = np.array([0, len(src)], np.int32), np.array([0, len(tgt)], np.int32)
    for kw in (dict(max_iter=1), dict(max_iter=3), dict(rel_fitness=1e-6, rel_rmse=1e-6, max_iter=30), dict(max_iter=0)):
        T_ref, ir = oicp.registration_icp_p2p(src, tgt, 10.0, relative_fitness=kw.get('rel_fitness', 1e-2),
                                              relative_rmse=kw.get('rel_rmse', 1e-2), max_iteration=kw['max_iter'], return_info=True)
        T, info = ops.icp_p2p(_dev(src), _dev(so), _dev(tgt), _dev(to), 10.0, **kw)
        assert int(info.cpu()[0, 2]) == ir['iterations']
        assert np.abs(T.cpu().numpy()[0] - T_ref).max() < 1e-5


def test_icp_large_grid_and_small_threshold():
    """Cloud much larger than the threshold: the cell size grows until the grid fits (<= 4096 cells)."""
    from autoposeestimation_b200 import ops
    rng = np.random.RandomState(5)
    tgt = rng.uniform(-300, 300, size=(6000, 3))
    R = synth.random_rotation(rng, math.radians(0.3)); t = np.array([0.4, -0.3, 0.2])
    src = (tgt[:2500] - t) @ R
    T_ref, ir = oicp.registration_icp_p2p(src, tgt, 2.0, return_info=True)
    T, info = ops.icp_p2p(_dev(src), _dev(np.array([0, 2500], np.int32)), _dev(tgt), _dev(np.array([0, 6000], np.int32)), 2.0)
    assert int(info.cpu()[0, 2]) == ir['iterations']
    assert np.abs(T.cpu().numpy()[0] - T_ref).max() < 1e-5


@pytest.mark.parametrize('box,threshold,n_t', [(60.0, 10.0, 1500), (60.0, 10.0, 60), (200.0, 10.0, 4000), (30.0, 3.0, 800)])
def test_icp_correspondence_search_exact(box, threshold, n_t):
    """The nearest-neighbour stage alone (max_iter=0: one evaluation of the correspondences) on clouds that stress the grid
    search: uniform targets in a box (nearest neighbours anywhere between 0 and the threshold, many queries with an empty own
    cell, queries outside the target's bounding box by less and by more than the threshold), for the half-threshold grid
    with a two-cell reach (60 mm and 30 mm boxes) and for the widened one-cell grid (200 mm box: too many cells).  The
    number of correspondences must equal the oracle's brute-force count exactly and the inlier rmse to 1e-12."""
    from autoposeestimation_b200 import ops
    rng = np.random.RandomState(int(box) + n_t)
    tgt = rng.uniform(0, box, size=(n_t, 3))
    src = rng.uniform(-1.5 * threshold, box + 1.5 * threshold, size=(3000, 3))
    src[:50] = tgt[:50] + rng.standard_normal((50, 3)) * 1e-3          # near-coincident points
    src[50:60] = tgt[50:60]                                            # exact hits (d = 0)
    idx, hit, fit, rmse = oicp._evaluate(src, tgt, threshold, oicp.nn_within)
    T, info = ops.icp_p2p(_dev(src), _dev(np.array([0, len(src)], np.int32)), _dev(tgt), _dev(np.array([0, n_t], np.int32)),
                          threshold, max_iter=0)
    info = info.cpu().numpy()[0]
    assert int(info[3]) == int(hit.sum()) and 0 < int(hit.sum()) < len(src)
    assert abs(info[0] - fit) < 1e-15 and abs(info[1] - rmse) < 1e-12
    assert np.array_equal(T.cpu().numpy()[0], np.identity(4))


@pytest.mark.parametrize('R', [592, 1776])
def test_icp_batch_property_identical_registrations(R):
    """Copies of one registration in a single launch give bit-identical transforms: 592 = one round at 4 CTAs per SM,
    1776 = the batch size at which the host switches to the 6-CTAs-per-SM build of the kernel (two rounds of 888)."""
    from autoposeestimation_b200 import ops
    src, tgt = _surface(4)
    S = np.tile(src, (R, 1)); Tg = np.tile(tgt, (R, 1))
    so = (np.arange(R + 1) * len(src)).astype(np.int32); to = (np.arange(R + 1) * len(tgt)).astype(np.int32)
    T, info = ops.icp_p2p(_dev(S), _dev(so), _dev(Tg), _dev(to), 10.0)
    assert bool((T == T[0:1]).all()) and bool((info == info[0:1]).all())
    T_ref = oicp.registration_icp_p2p(src, tgt, 10.0)
    assert np.abs(T[0].cpu().numpy() - T_ref).max() < 1e-5


def test_icp_wide_cta_matches_narrow_kernel():
    """Launches of at most one registration per SM run a 512-thread CTA per registration (the sequential reconstruction
    loop registers one view at a time); the same registrations inside a batch larger than the SM count run the
    128-thread kernel: same iteration and correspondence counts, transforms equal to 1e-9 (the fixed reduction trees
    differ), both inside the parity gate against the oracle."""
    from autoposeestimation_b200 import ops
    probs = [_surface(s) for s in (0, 3, 5)]
    S, so = _ragged([p[0] for p in probs]); Tg, to = _ragged([p[1] for p in probs])
    Tw, iw = ops.icp_p2p(_dev(S), _dev(so), _dev(Tg), _dev(to), 10.0)                      # 3 registrations: wide CTAs
    reps = 60                                                                                # 180 registrations: 128-thread kernel
    Sn, son = _ragged([p[0] for p in probs] * reps); Tn_, ton = _ragged([p[1] for p in probs] * reps)
    Tn, inn = ops.icp_p2p(_dev(Sn), _dev(son), _dev(Tn_), _dev(ton), 10.0)
    Tw, iw, Tn, inn = (x.cpu().numpy() for x in (Tw, iw, Tn, inn))
    for k, (src, tgt) in enumerate(probs):
        assert np.array_equal(iw[k, 2:], inn[k, 2:]) and np.array_equal(inn[k], inn[k + 3 * (reps - 1)])
        assert np.abs(Tw[k] - Tn[k]).max() < 1e-9 and abs(iw[k, 1] - inn[k, 1]) < 1e-9
        assert np.abs(Tw[k] - oicp.registration_icp_p2p(src, tgt, 10.0)).max() < 1e-5


@pytest.mark.parametrize('voxel', [2.0, 5.0])
def test_voxel_down_sample_exact(voxel):
    from autoposeestimation_b200 import ops
    rng = np.random.RandomState(int(voxel))
    clouds = [synth.ellipsoid_cloud(rng, 5000) + rng.standard_normal((5000, 3)), rng.uniform(-50, 50, size=(16384, 3)),
              rng.uniform(0, 1, size=(3, 3)), np.zeros((0, 3)), np.array([[1.0, 2.0, 3.0]])]
    P, off = _ragged(clouds)
    out, cnt = ops.voxel_down_sample(_dev(P), _dev(off), voxel)
    out = out.cpu().numpy(); cnt = cnt.cpu().numpy()
    for c, cl in enumerate(clouds):
        want = oicp.voxel_down_sample(cl, voxel)
        assert cnt[c] == len(want)
        assert np.array_equal(out[off[c]:off[c] + cnt[c]], want)               # same order, same bits as the oracle


def test_voxel_big_batch_mixed_sizes():
    """A batch of several hundred clouds of mixed sizes (empty, 1 point, around the powers of two the in-CTA sort pads to,
    the 16 384-point limit and one cloud above it, flagged -1): every cloud's output has the bits it has in a small batch
    and, for sampled clouds, the oracle's."""
    from autoposeestimation_b200 import ops
    rng = np.random.RandomState(11)
    sizes = [0, 1, 4095, 4096, 4097, 8191, 8192, 8193, 16384, 16385] + list(rng.randint(1, 12000, size=330))
    clouds = [rng.uniform(-40, 40, size=(n, 3)) for n in sizes]
    P, off = _ragged(clouds)
    out, cnt = ops.voxel_down_sample(_dev(P), _dev(off), 3.0, max_cloud_points=16384)       # 340 clouds in one launch
    out = out.cpu().numpy(); cnt = cnt.cpu().numpy()
    assert cnt[0] == 0 and cnt[9] == -1
    for lo in range(0, len(clouds), 100):                                                   # the same clouds, 100 per launch
        sub = clouds[lo:lo + 100]
        Ps, offs = _ragged(sub)
        o2, c2 = ops.voxel_down_sample(_dev(Ps), _dev(offs), 3.0, max_cloud_points=16384)
        o2 = o2.cpu().numpy(); c2 = c2.cpu().numpy()
        for k in range(len(sub)):
            assert c2[k] == cnt[lo + k]
            if c2[k] > 0:
                assert np.array_equal(o2[offs[k]:offs[k] + c2[k]], out[off[lo + k]:off[lo + k] + c2[k]])
    for k in (2, 4, 7, 8):
        assert np.array_equal(out[off[k]:off[k] + cnt[k]], oicp.voxel_down_sample(clouds[k], 3.0))


def test_voxel_batched_entry_flags_oversize_clouds():
    """The batched C entry itself does not process a cloud above 16 384 points: it flags it with out_counts = -1 (header
    contract) and ops.voxel_down_sample then routes it through the large path; when the caller vouches for the sizes
    (max_cloud_points) the flag is what comes back."""
    from autoposeestimation_b200 import ops
    P = torch.rand((16385, 3), dtype=torch.float64, device='cuda')
    off = torch.tensor([0, 16385], dtype=torch.int32, device='cuda')
    out, cnt = ops.voxel_down_sample(P, off, 0.01, max_cloud_points=16384)
    assert int(cnt[0]) == -1
    out, cnt = ops.voxel_down_sample(P, off, 0.01)
    assert int(cnt[0]) == len(oicp.voxel_down_sample(P.cpu().numpy(), 0.01))


@pytest.mark.parametrize('n', [16385, 40000, 307200])
def test_voxel_down_sample_large_clouds(n):
    """Clouds beyond the 16 384-point shared-memory sort (a full 480 x 640 mask has 307 200 pixels): the global-memory path
    (chunk sort + merge passes) gives the oracle's result bit for bit, alone and inside a ragged batch with small clouds."""
    from autoposeestimation_b200 import ops
    rng = np.random.RandomState(n)
    big = rng.uniform(-150, 150, size=(n, 3)) * np.array([1.0, 0.6, 0.3])
    small = rng.uniform(-40, 40, size=(3000, 3))
    want_big, want_small = oicp.voxel_down_sample(big, 2.0), oicp.voxel_down_sample(small, 2.0)
    pts, off = _ragged([small, big, small[:7]])
    out, cnt = ops.voxel_down_sample(_dev(pts), _dev(off), 2.0, offset_host=off)
    cnt = cnt.cpu().numpy(); out = out.cpu().numpy()
    assert cnt[0] == len(want_small) and np.array_equal(out[off[0]:off[0] + cnt[0]], want_small)
    assert cnt[1] == len(want_big) and np.array_equal(out[off[1]:off[1] + cnt[1]], want_big)
    assert cnt[2] == len(oicp.voxel_down_sample(small[:7], 2.0))
    out2, cnt2 = ops.voxel_down_sample(_dev(pts), _dev(off), 2.0)               # sizes fetched from the device
    assert np.array_equal(cnt2.cpu().numpy(), cnt) and np.array_equal(out2.cpu().numpy()[off[1]:off[1] + cnt[1]], want_big)
    from autoposeestimation_b200.pc_reconstruction.open3d_utils import PointCloud
    pc = PointCloud(big).voxel_down_sample(5.0)                                  # the drop-in class no longer has a size limit
    assert np.array_equal(pc.numpy(), oicp.voxel_down_sample(big, 5.0))
