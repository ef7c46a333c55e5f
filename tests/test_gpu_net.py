"""GPU parity: PoseNet geometry + PoseRefineNet + pose pipeline vs the torch-CPU oracle and the golden
vectors of the reference's own modules.  Gates (BASELINE.json): translation 1e-4 m, rotation 1e-3 rad."""
import os

import numpy as np
import pytest
import torch

from oracle import densefusion as odf, pose_math as pm, synth

pytestmark = pytest.mark.gpu
TOL_T, TOL_R = 1e-4, 1e-3


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _handles(seed, nobj, max_batch, max_points):
    from autoposeestimation_b200 import ops
    sd_e = synth.posenet_state_dict(seed, nobj); sd_r = synth.refiner_state_dict(seed + 1000, nobj)
    est = ops.NetHandle(ops.NET_POSENET, sd_e, nobj, max_batch, max_points)
    ref = ops.NetHandle(ops.NET_REFINER, sd_r, nobj, max_batch, max_points)
    return est, ref, synth.to_torch(sd_e), synth.to_torch(sd_r)


@pytest.mark.parametrize('impl', ['simt', 'tcgen05', 'tcgen05_v1', 'tcgen05_pair', 'tcgen05_b2b'])
@pytest.mark.parametrize('case', [0, 1, 2])
def test_posenet_refiner_golden(golden_dir, case, impl):
    """Raw network outputs vs the reference's own PoseNet / PoseRefineNet (tests/golden)."""
    from autoposeestimation_b200 import ops
    g = np.load(os.path.join(golden_dir, 'densefusion_case%d.npz' % case))
    seed, npts, nobj = int(g['seed']), int(g['npts']), int(g['nobj'])
    hw = tuple(int(v) for v in g['hw'])
    est, ref, _, _ = _handles(seed, nobj, 2, npts)
    gi = {'simt': ops.GEMM_SIMT, 'tcgen05': ops.GEMM_TCGEN05, 'tcgen05_v1': ops.GEMM_TCGEN05_V1, 'tcgen05_pair': ops.GEMM_TCGEN05_PAIR,
          'tcgen05_b2b': ops.GEMM_TCGEN05_B2B}[impl]
    est.set_gemm(gi); ref.set_gemm(gi)
    out_img, cloud, choose, idx = synth.posenet_inputs(seed, npts, hw, nobj)
    r, t, c, emb = est.posenet_forward(_dev(out_img), _dev(cloud), _dev(choose), _dev(idx))
    torch.cuda.synchronize()
    scale_r = np.abs(g['r']).max(); scale_t = np.abs(g['t']).max()
    assert np.abs(r.cpu().numpy() - g['r']).max() < 2e-4 * max(1.0, scale_r)
    assert np.abs(t.cpu().numpy() - g['t']).max() < 2e-4 * max(1.0, scale_t)
    assert np.abs(c.cpu().numpy() - g['c']).max() < 1e-4
    assert np.allclose(emb.cpu().numpy().sum(axis=1), g['emb_sum'], atol=1e-4)
    r2, t2 = ref.refiner_forward(_dev(g['new_points']), emb, _dev(idx))
    assert np.abs(r2.cpu().numpy() - g['r2']).max() < 2e-4 * max(1.0, np.abs(g['r2']).max())
    assert np.abs(t2.cpu().numpy() - g['t2']).max() < 2e-4 * max(1.0, np.abs(g['t2']).max())


def _check_pose(pose, want_q, want_t, c_sorted_gap, tag):
    ang = pm.rotation_angle_between(pose[:4], want_q)
    dt = np.abs(pose[4:] - want_t).max()
    ok = ang < TOL_R and dt < TOL_T
    # An arg-max flip is tolerated only for a confidence near-tie (SURVEY 7 "hard parts")
    assert ok or c_sorted_gap < 1e-5, (tag, ang, dt, c_sorted_gap)
    return ok


@pytest.mark.parametrize('canonical', [False, True])
def test_pose_pipeline_vs_oracle(canonical):
    """Batched pipeline (B=6 objects, N=500) vs the per-sample reference loops restated in the oracle."""
    from autoposeestimation_b200 import ops
    nobj, B, N = 5, 6, 500
    est, ref, sd_e, sd_r = _handles(31, nobj, B, N)
    out_img, cloud, choose, idx = synth.posenet_inputs(32, N, (120, 160), nobj, batch=B)
    poses, wm = ops.pose_pipeline(est, ref, _dev(out_img), _dev(cloud), _dev(choose), _dev(idx), iterations=2, canonical=canonical)
    poses = poses.cpu().numpy(); wm = wm.cpu().numpy()
    torch.set_num_threads(4)
    n_ok = 0
    for b in range(B):
        args = (sd_e, sd_r, torch.from_numpy(out_img[b:b + 1]), torch.from_numpy(cloud[b:b + 1]),
                torch.from_numpy(choose[b:b + 1]), torch.from_numpy(idx[b:b + 1]), nobj)
        with torch.no_grad():
            res = odf.canonical_prediction(*args, iterations=2) if canonical else odf.live_prediction(*args, refine_calls=2)
            _, _, c, _ = odf.posenet_geometry(sd_e, *args[2:6], nobj)
        cs = np.sort(c.numpy().reshape(-1))
        n_ok += _check_pose(poses[b], res['q'], res['t'], cs[-1] - cs[-2], 'obj %d' % b)
        if cs[-1] - cs[-2] > 1e-5:
            assert wm[b] == res['which_max']
    assert n_ok >= B - 1


def test_pose_pipeline_no_refine_and_n1000():
    from autoposeestimation_b200 import ops
    nobj, B, N = 3, 2, 1000
    est, ref, sd_e, sd_r = _handles(41, nobj, B, N)
    out_img, cloud, choose, idx = synth.posenet_inputs(42, N, (80, 120), nobj, batch=B)
    poses, wm = ops.pose_pipeline(est, None, _dev(out_img), _dev(cloud), _dev(choose), _dev(idx), iterations=0)
    poses = poses.cpu().numpy()
    for b in range(B):
        with torch.no_grad():
            res = odf.live_prediction(sd_e, sd_r, torch.from_numpy(out_img[b:b + 1]), torch.from_numpy(cloud[b:b + 1]),
                                      torch.from_numpy(choose[b:b + 1]), torch.from_numpy(idx[b:b + 1]), nobj, refine_calls=0)
        assert pm.rotation_angle_between(poses[b, :4], res['q']) < TOL_R and np.abs(poses[b, 4:] - res['t']).max() < TOL_T


@pytest.mark.parametrize('B,N', [(64, 500), (130, 256), (17, 300)])
def test_tcgen05_matches_simt_batch64(B, N):
    """BASELINE config 2 shape (batch 64, N=500) and two more batch sizes (130: 256-column dense tiles; 17: the smallest batch
    on the tensor-core dense layers): the tcgen05 path (GEMM layers + per-object dense layers, gemm_dense.cuh) against the
    SIMT fp32 kernels on the same buffers."""
    from autoposeestimation_b200 import ops
    nobj = 5
    est, ref, _, _ = _handles(51, nobj, B, N)
    out_img, cloud, choose, idx = synth.posenet_inputs(52, N, (120, 160), nobj, batch=B)
    d = [_dev(a) for a in (out_img, cloud, choose, idx)]
    outs = {}
    for name, gi in (('simt', ops.GEMM_SIMT), ('tc', ops.GEMM_TCGEN05), ('b2b', ops.GEMM_TCGEN05_B2B)):
        est.set_gemm(gi); ref.set_gemm(gi)
        outs[name] = [x.clone() for x in est.posenet_forward(*d)]
        poses, _ = ops.pose_pipeline(est, ref, *d, iterations=2, canonical=True)
        outs[name].append(poses.clone())
    for other in ('tc', 'b2b'):                                  # b2b: conv1 -> conv2 of the heads fused back to back (gemm_tc4.cuh)
        for a, b in zip(outs['simt'][:3], outs[other][:3]):
            assert float((a - b).abs().max()) < 5e-5 * max(1.0, float(a.abs().max()))
        dq = (outs['simt'][4][:, :4] - outs[other][4][:, :4]).abs().max(dim=1).values
        dt = (outs['simt'][4][:, 4:] - outs[other][4][:, 4:]).abs().max(dim=1).values
        assert int(((dq < 1e-4) & (dt < 1e-4)).sum()) >= B - 1  # at most one arg-max near-tie flip


def test_config2_batch64_sampled_objects_vs_oracle():
    """BASELINE config 2 AT SCALE (batch 64, N=500, 2 canonical refine iterations) against the oracle itself: objects
    sampled from different 128-row tiles of the batch are re-run through the per-sample fp32 restatement of the reference
    (pinned by the reference-generated golden vectors)."""
    from autoposeestimation_b200 import ops
    nobj, B, N = 5, 64, 500
    est, ref, sd_e, sd_r = _handles(71, nobj, B, N)
    out_img, cloud, choose, idx = synth.posenet_inputs(72, N, (120, 160), nobj, batch=B)
    poses, wm = ops.pose_pipeline(est, ref, _dev(out_img), _dev(cloud), _dev(choose), _dev(idx), iterations=2, canonical=True)
    poses = poses.cpu().numpy(); wm = wm.cpu().numpy()
    torch.set_num_threads(4)
    sample = [0, 17, 38, 63]
    n_ok = 0
    for b in sample:
        args = (sd_e, sd_r, torch.from_numpy(out_img[b:b + 1]), torch.from_numpy(cloud[b:b + 1]),
                torch.from_numpy(choose[b:b + 1]), torch.from_numpy(idx[b:b + 1]), nobj)
        with torch.no_grad():
            res = odf.canonical_prediction(*args, iterations=2)
            _, _, c, _ = odf.posenet_geometry(sd_e, *args[2:6], nobj)
        cs = np.sort(c.numpy().reshape(-1))
        n_ok += _check_pose(poses[b], res['q'], res['t'], cs[-1] - cs[-2], 'obj %d' % b)
        if cs[-1] - cs[-2] > 1e-5:
            assert wm[b] == res['which_max']
    assert n_ok >= len(sample) - 1


def test_split_bf16_pass_table():
    """Per-layer product table (ape_net_set_passes): the default drops A_lo*W_hi only in the layers whose output is pooled
    over the points; against all-three-products everywhere the poses move by far less than the gate, clouds below 256
    points always get all three products (bit-identical), and a mask without A_hi*W_hi is rejected."""
    from autoposeestimation_b200 import ops, _lib
    nobj, B, N = 5, 8, 500
    est, ref, _, _ = _handles(81, nobj, B, N)
    assert est.get_passes() == [7, 6, 6, 7, 7, 7] and ref.get_passes()[:3] == [6, 6, 6]
    d = [_dev(a) for a in synth.posenet_inputs(82, N, (120, 160), nobj, batch=B)]
    p_def = ops.pose_pipeline(est, ref, *d)[0].cpu().numpy()
    est.set_passes([7] * 6); ref.set_passes([7] * 6)
    p_full = ops.pose_pipeline(est, ref, *d)[0].cpu().numpy()
    for b in range(B):
        assert pm.rotation_angle_between(p_def[b, :4], p_full[b, :4]) < 2e-4 and np.abs(p_def[b, 4:] - p_full[b, 4:]).max() < 2e-5
    assert not np.array_equal(p_def, p_full)                      # the table is really in effect
    small = [_dev(a) for a in synth.posenet_inputs(83, 128, (40, 40), nobj, batch=B)]
    p_small_full = ops.pose_pipeline(est, ref, *small)[0].clone()
    est.set_passes([7, 6, 6, 7, 7, 7]); ref.set_passes([6, 6, 6])
    assert torch.equal(ops.pose_pipeline(est, ref, *small)[0], p_small_full)
    with pytest.raises(_lib.ApeError):
        est.set_passes([3, 7, 7, 7, 7, 7])


def test_net_errors_are_loud():
    from autoposeestimation_b200 import ops, _lib
    est, ref, _, _ = _handles(61, 2, 1, 128)
    out_img, cloud, choose, idx = synth.posenet_inputs(62, 256, (40, 40), 2)
    with pytest.raises(_lib.ApeError):
        est.posenet_forward(_dev(out_img), _dev(cloud), _dev(choose), _dev(idx))      # N exceeds the workspace
    with pytest.raises(_lib.ApeError):
        ref.posenet_forward(_dev(out_img[:, :, :4, :32]), _dev(cloud[:, :128]), _dev(choose[:, :, :128]), _dev(idx))   # wrong kind


_TAIL_WORKER = r'''
import sys, numpy as np, torch
sys.path.insert(0, sys.argv[1])
from autoposeestimation_b200 import ops
from oracle import synth
out = {}
for B, N in ((1, 500), (5, 300), (64, 500), (65, 256), (200, 128)):
    nobj = 4
    ref = ops.NetHandle(ops.NET_REFINER, synth.refiner_state_dict(300 + B, nobj), nobj, B, N)
    rng = np.random.RandomState(B)
    pts = torch.from_numpy((rng.randn(B, N, 3) * 0.05).astype(np.float32)).cuda()
    emb = torch.from_numpy(rng.randn(B, 32, N).astype(np.float32)).cuda()
    idx = torch.from_numpy(rng.randint(0, nobj, (B,)).astype(np.int64)).cuda()
    r2, t2 = ref.refiner_forward(pts, emb, idx)
    r2b, t2b = ref.refiner_forward(pts, emb, idx)
    assert torch.equal(r2, r2b) and torch.equal(t2, t2b)          # deterministic (fixed reduction order)
    out['r%d' % B] = r2.cpu().numpy(); out['t%d' % B] = t2.cpu().numpy()
np.savez(sys.argv[2], **out)
'''


def test_refiner_tail_cluster_kernel_vs_separate_launches_and_oracle(tmp_path):
    """The one-launch PoseRefineNet tail (gemm_tail.cuh: pooling finish + conv1..conv3 in a 16- or 8-CTA cluster) against
    the separate pool_finish / dense / refiner_out launches (APE_REFINER_TAIL=0) and against the oracle's fp32 refiner,
    at batch sizes around the 64-object chunk (1, 5, 64, 65, 200); r2 / t2 are deterministic run to run."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / 'tail_worker.py'
    script.write_text(_TAIL_WORKER)
    res = {}
    for mode in ('0', '8', '16'):
        env = dict(os.environ, APE_REFINER_TAIL=mode, PYTHONPATH=root)
        out = tmp_path / ('tail_%s.npz' % mode)
        r = subprocess.run([sys.executable, str(script), root, str(out)], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        res[mode] = np.load(out)
    for key in res['0'].files:
        a = res['0'][key]
        for mode in ('8', '16'):
            assert np.abs(res[mode][key] - a).max() < 2e-5 * max(1.0, np.abs(a).max()), (key, mode)
    # and the oracle (per-sample fp32 restatement of network.py:187-204) on a few objects of the 65-object case
    B, N, nobj = 65, 256, 4
    sd = synth.to_torch(synth.refiner_state_dict(300 + B, nobj))
    rng = np.random.RandomState(B)
    pts = (rng.randn(B, N, 3) * 0.05).astype(np.float32); emb = rng.randn(B, 32, N).astype(np.float32)
    idx = rng.randint(0, nobj, (B,)).astype(np.int64)
    for b in (0, 63, 64):
        r, t = odf.refiner_forward(sd, torch.from_numpy(pts[b:b + 1]), torch.from_numpy(emb[b:b + 1]), torch.from_numpy(idx[b:b + 1]).view(1, 1), nobj)
        assert np.abs(res['16']['r65'][b] - r[0].detach().numpy()).max() < 2e-4
        assert np.abs(res['16']['t65'][b] - t[0].detach().numpy()).max() < 2e-4
