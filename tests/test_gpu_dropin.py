"""GPU: the reference-facing drop-in modules (same names / signatures as the reference) against golden vectors of
the reference's own modules and against the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import densefusion as odf, geometry as og, icp as oicp, pose_math as pm, synth

pytestmark = pytest.mark.gpu


def _modules(seed, npts, nobj):
    from autoposeestimation_b200.densefusion import network
    est = network.PoseNet(npts, nobj); ref = network.PoseRefineNet(npts, nobj)
    est.load_state_dict(synth.to_torch(synth.posenet_state_dict(seed, nobj)), strict=False)
    ref.load_state_dict(synth.to_torch(synth.refiner_state_dict(seed + 1000, nobj)), strict=True)
    return est.cuda().eval(), ref.cuda().eval()


@pytest.mark.parametrize('case', [0, 1])
def test_modules_reproduce_reference_flow(golden_dir, case):
    """estimator -> get_new_points -> my_estimator_prediction -> refiner -> my_refined_prediction, exactly as
    pipeline/utils.py:564-571 calls them, against the reference's own outputs (tests/golden)."""
    from autoposeestimation_b200.densefusion.tools_utils import get_new_points, my_estimator_prediction, my_refined_prediction
    g = np.load(os.path.join(golden_dir, 'densefusion_case%d.npz' % case))
    seed, npts, nobj = int(g['seed']), int(g['npts']), int(g['nobj'])
    hw = tuple(int(v) for v in g['hw'])
    est, ref = _modules(seed, npts, nobj)
    est.cnn = torch.nn.Identity()                      # `img` is the encoder output, as in oracle/gen_golden.py
    out_img, cloud, choose, idx = (torch.from_numpy(a).cuda() for a in synth.posenet_inputs(seed, npts, hw, nobj))
    with torch.no_grad():
        pred_r, pred_t, pred_c, emb = est(out_img, cloud, choose, idx)
        assert pred_r.shape == (1, npts, 4) and pred_t.shape == (1, npts, 3) and pred_c.shape == (1, npts, 1) and emb.shape == (1, 32, npts)
        new_points = get_new_points(pred_r, pred_t, pred_c, cloud)
        _, my_r, my_t = my_estimator_prediction(pred_r, pred_t, pred_c, npts, 1, cloud)
        for _ in range(2):
            r2, t2 = ref(new_points, emb, idx)
        my_pred, fq, ft = my_refined_prediction(r2, t2, my_r, my_t)
    assert np.allclose(new_points.cpu().numpy(), g['new_points'], atol=2e-5)
    assert np.allclose(my_r, g['my_r'], atol=2e-5) and np.allclose(my_t, g['my_t'], atol=2e-5)
    assert pm.rotation_angle_between(fq, g['final_q']) < 1e-3 and np.abs(ft - g['final_t']).max() < 1e-4
    assert my_pred.shape == (7,)


def test_module_handle_refreshes_after_weight_update():
    est, ref = _modules(3, 128, 2)
    x = torch.randn(1, 128, 3, device='cuda'); emb = torch.randn(1, 32, 128, device='cuda'); idx = torch.zeros(1, 1, dtype=torch.long, device='cuda')
    with torch.no_grad():
        a = ref(x, emb, idx)[1].clone()
        ref.conv3_t.bias.add_(1.0)                     # in-place parameter update (optimizer step) -> version bump
        b = ref(x, emb, idx)[1]
    assert torch.allclose(b - a, torch.ones_like(a), atol=1e-5)


def test_knearestneighbor_and_loss_refine_symmetric(golden_dir):
    from autoposeestimation_b200.densefusion.knn import KNearestNeighbor
    from autoposeestimation_b200.densefusion.loss import Loss
    from autoposeestimation_b200.densefusion.loss_refiner import Loss_refine
    g = np.load(os.path.join(golden_dir, 'knn.npz'))
    knn = KNearestNeighbor(1)
    inds = knn(torch.from_numpy(g['ref']), torch.from_numpy(g['qry']))         # CPU tensors are moved, as the reference does
    assert inds.is_cuda and inds.dtype == torch.int64 and np.array_equal(inds.cpu().numpy(), g['idx_k1'])
    gl = np.load(os.path.join(golden_dir, 'losses.npz'))
    T = lambda k: torch.from_numpy(gl[k]).cuda()
    idx = torch.zeros((1, 1), dtype=torch.long, device='cuda')
    pr = T('pr1').requires_grad_(True)
    dis, npn, ntg, pred = Loss_refine(120, [0])(pr, T('pt1'), T('target'), T('model'), idx, T('points'))
    assert abs(float(dis) - float(gl['lr_dis_sym'][0])) < 1e-6                   # ADD-S gate 1e-5 m
    assert np.allclose(npn.cpu().numpy(), gl['lr_newp_sym'], atol=1e-6) and np.allclose(pred.detach().cpu().numpy(), gl['lr_pred_sym'], atol=1e-6)
    dis.backward()
    assert pr.grad is not None and float(pr.grad.abs().sum()) > 0                # differentiable (train.py:222)
    lo, d, npn, ntg, _ = Loss(120, [0])(T('pr_n'), T('pt_n'), T('pc_n'), T('target'), T('model'), idx, T('points'), 0.015, False)
    assert abs(float(lo) - float(gl['l_loss_sym'])) < 1e-5 and abs(float(d) - float(gl['l_dis_sym'])) < 1e-6


def test_loss_dropins_nonsymmetric_match_reference(golden_dir):
    """Loss / Loss_refine (non-symmetric branch) vs the outputs of the reference's own modules (tests/golden/losses.npz)."""
    from autoposeestimation_b200.densefusion.loss import Loss
    from autoposeestimation_b200.densefusion.loss_refiner import Loss_refine
    g = np.load(os.path.join(golden_dir, 'losses.npz'))
    T = lambda k: torch.from_numpy(g[k]).cuda()
    idx = torch.zeros((1, 1), dtype=torch.long, device='cuda')
    dis, npn, ntg, pred = Loss_refine(120, [])(T('pr1'), T('pt1'), T('target'), T('model'), idx, T('points'))
    N = lambda x: x.detach().cpu().numpy()
    assert np.allclose(N(dis), g['lr_dis_nosym'], atol=1e-6) and np.allclose(N(npn), g['lr_newp_nosym'], atol=1e-6)
    assert np.allclose(N(ntg), g['lr_newt_nosym'], atol=1e-6) and np.allclose(N(pred), g['lr_pred_nosym'], atol=1e-6)
    for tag, sym, refine in (('nosym', [], False), ('symrefine', [0], True)):
        lo, d, npn, ntg, _ = Loss(120, sym)(T('pr_n'), T('pt_n'), T('pc_n'), T('target'), T('model'), idx, T('points'), 0.015, refine)
        assert np.allclose(N(lo), g['l_loss_' + tag], atol=1e-6) and np.allclose(N(d), g['l_dis_' + tag], atol=1e-6)
        assert np.allclose(N(npn), g['l_newp_' + tag], atol=1e-6) and np.allclose(N(ntg), g['l_newt_' + tag], atol=1e-6)


def test_loss_forward_and_gradient_match_reference_autograd(golden_dir):
    """`Loss` on the kernels (csrc/knn.cu: estimator_loss_kernel) against the reference's own lib/loss.py differentiated by
    torch autograd (tests/golden/loss_grads.npz, oracle/gen_golden.py): loss, `pred`, and d loss / d (pred_r, pred_t, pred_c)
    for a symmetric class (nearest-neighbour targets, indices detached) and a non-symmetric one; a no-grad call returns
    the same values."""
    from autoposeestimation_b200.densefusion.loss import Loss
    g = np.load(os.path.join(golden_dir, 'losses.npz'))
    gg = np.load(os.path.join(golden_dir, 'loss_grads.npz'))
    T = lambda k: torch.from_numpy(g[k]).cuda()
    idx = torch.zeros((1, 1), dtype=torch.long, device='cuda')
    for tag, sym in (('sym', [0]), ('nosym', [])):
        pr, pt, pc = T('pr_n').requires_grad_(True), T('pt_n').requires_grad_(True), T('pc_n').requires_grad_(True)
        lo, dis, npn, ntg, pred = Loss(120, sym)(pr, pt, pc, T('target'), T('model'), idx, T('points'), 0.015, False)
        assert abs(float(lo) - float(gg['loss_' + tag])) < 1e-6
        assert np.allclose(pred.cpu().numpy(), gg['pred_' + tag], atol=1e-6)
        assert np.allclose(npn.cpu().numpy(), g['l_newp_' + tag], atol=1e-6) and np.allclose(ntg.cpu().numpy(), g['l_newt_' + tag], atol=1e-6)
        assert abs(float(dis) - float(g['l_dis_' + tag])) < 1e-6 and not dis.requires_grad and not npn.requires_grad
        (2.0 * lo).backward()
        for got, key in ((pr.grad, 'd_r_'), (pt.grad, 'd_t_'), (pc.grad, 'd_c_')):
            want = 2.0 * gg[key + tag]
            assert got.shape == want.shape
            assert np.allclose(got.cpu().numpy(), want, rtol=2e-4, atol=1e-7), (tag, key, np.abs(got.cpu().numpy() - want).max())
        with torch.no_grad():
            lo2, dis2, *_ = Loss(120, sym)(T('pr_n'), T('pt_n'), T('pc_n'), T('target'), T('model'), idx, T('points'), 0.015, False)
        assert float(lo2) == float(lo) and float(dis2) == float(dis)
    with pytest.raises(NotImplementedError):
        Loss(120, [])(T('pr_n').repeat(2, 1, 1), T('pt_n').repeat(2, 1, 1), T('pc_n').repeat(2, 1, 1), T('target').repeat(2, 1, 1),
                      T('model').repeat(2, 1, 1), idx, T('points').repeat(2, 1, 1), 0.015, False)


def test_loss_large_mesh_and_many_candidates():
    """Shapes of the reference's training configs: 1000 candidates x 1000 mesh points (this fork) and a 2600-point mesh (YCB
    refine phase), symmetric: per-candidate dis / std against the ADD-S kernel pinned elsewhere, finite gradients."""
    from autoposeestimation_b200 import ops
    rng = np.random.RandomState(4)
    for N, M in ((1000, 1000), (300, 2600)):
        model = torch.from_numpy(rng.uniform(-0.1, 0.1, (M, 3)).astype(np.float32)).cuda()
        target = model + 0.003 * torch.randn_like(model)
        pr = torch.from_numpy(rng.standard_normal((N, 4)).astype(np.float32)).cuda(); pr[:, 0] += 3.0
        pt = torch.from_numpy((rng.standard_normal((N, 3)) * 0.01).astype(np.float32)).cuda()
        pc = torch.from_numpy(rng.uniform(0.1, 0.9, N).astype(np.float32)).cuda()
        pts = torch.from_numpy((rng.standard_normal((N, 3)) * 0.01).astype(np.float32)).cuda()
        out = ops.estimator_loss(pr, pt, pc, pts, model, target, True, 0.015)
        dis = ops.add_metric(pr, pts + pt, model, target, torch.ones(N, dtype=torch.uint8, device='cuda'))
        assert torch.allclose(out['dis'], dis, atol=1e-6)
        want = torch.mean((out['dis'] + 2 * out['std']) * pc - 0.015 * torch.log(pc))
        assert abs(float(out['loss']) - float(want)) < 1e-6
        assert int(out['which_max'][0]) == int(torch.argmax(pc))
        for k in ('d_r', 'd_t', 'd_c'):
            assert bool(torch.isfinite(out[k]).all()) and float(out[k].abs().sum()) > 0


def test_refiner_dropin_trains_like_reference_loop(golden_dir):
    """train.py:215-233 written against the drop-ins: refiner.train(); per sample and iteration
    `refiner(new_points, emb, idx)` -> `Loss_refine` -> `dis.backward()`; `optim.Adam(refiner.parameters()).step()`.
    The gradients land in the parameters' .grad (views of one flat vector) and equal the batched trainer's."""
    from autoposeestimation_b200 import synthetic as synth
    from autoposeestimation_b200.densefusion import network
    from autoposeestimation_b200.densefusion.loss_refiner import Loss_refine
    from autoposeestimation_b200.densefusion.train_refiner import RefinerTrainer
    nobj, N, M = 3, 200, 150
    sd = synth.refiner_state_dict(77, nobj)
    sd['conv3_r.bias'] = sd['conv3_r.bias'].copy(); sd['conv3_r.bias'][0::4] += 1.0
    rng = np.random.RandomState(5)
    B = 3
    pts = torch.from_numpy((rng.randn(B, N, 3) * 0.05).astype(np.float32)).cuda()
    emb = torch.from_numpy(rng.randn(B, 32, N).astype(np.float32)).cuda()
    idx = torch.from_numpy(rng.randint(0, nobj, (B, 1)).astype(np.int64)).cuda()
    model = torch.from_numpy(((rng.rand(B, M, 3) - 0.5) * 0.2).astype(np.float32)).cuda()
    target = model + 0.01
    refiner = network.PoseRefineNet(N, nobj).cuda()
    refiner.load_state_dict(synth.to_torch(sd))
    refiner.train()
    optimizer = torch.optim.Adam(refiner.parameters(), lr=1e-4)
    criterion_refine = Loss_refine(M, [1])
    optimizer.zero_grad()
    for b in range(B):                                               # the reference's batch-1 loop
        p, tg = pts[b:b + 1], target[b:b + 1]
        for ite in range(2):
            pred_r, pred_t = refiner(p, emb[b:b + 1], idx[b:b + 1])
            dis, p, tg, _ = criterion_refine(pred_r, pred_t, tg, model[b:b + 1], idx[b:b + 1], p)
            dis.backward()
    trainer = RefinerTrainer(sd, nobj, B, N, sym_list=[1])
    trainer.zero_grad()
    trainer.accumulate(pts, emb, idx.view(-1), target, model)
    flat = refiner.flat_gradient()
    assert refiner.conv1_r.weight.grad.data_ptr() == refiner._tr.view('conv1_r.weight', flat).data_ptr()
    rel = float((flat - trainer.h.grads).norm() / trainer.h.grads.norm())
    assert rel < 2e-3, rel                                            # same kernels; only fp32 atomic order / batching differ
    before = refiner.conv1_r.weight.detach().clone()
    r_before = refiner(pts, emb, idx)[0].detach().clone()
    optimizer.step()
    assert float((refiner.conv1_r.weight - before).abs().max()) > 0
    refiner.eval()
    with torch.no_grad():
        r_after = refiner(pts, emb, idx)[0]
    assert float((r_after - r_before).abs().max()) > 0               # the bf16 weight copies follow the in-place update


def test_refiner_autograd_patterns_other_than_the_reference_loop():
    """Patterns autograd supports but train.py does not use: a loss summed over two forwards before ONE backward, and an
    eval / no_grad forward between a training forward and its backward.  Both must give the gradient of the plain
    forward -> backward sequence (the backward re-runs its own forward when the workspace was overwritten)."""
    from autoposeestimation_b200 import synthetic as synth
    from autoposeestimation_b200.densefusion import network
    nobj, N = 3, 200
    sd = synth.refiner_state_dict(91, nobj)
    rng = np.random.RandomState(6)
    pts = [torch.from_numpy((rng.randn(1, N, 3) * 0.05).astype(np.float32)).cuda() for _ in range(2)]
    emb = [torch.from_numpy(rng.randn(1, 32, N).astype(np.float32)).cuda() for _ in range(2)]
    idx = torch.zeros((1, 1), dtype=torch.long, device='cuda')
    w = [torch.from_numpy(rng.randn(7).astype(np.float32)).cuda() for _ in range(2)]

    def loss_of(refiner, i):
        r, t = refiner(pts[i], emb[i], idx)
        return (torch.cat([r.view(-1), t.view(-1)]) * w[i]).sum()

    def fresh():
        m = network.PoseRefineNet(N, nobj).cuda(); m.load_state_dict(synth.to_torch(sd)); m.train()
        return m
    a = fresh()
    loss_of(a, 0).backward(); loss_of(a, 1).backward()                # the reference's pattern
    want = a.flat_gradient().clone()
    b = fresh()
    (loss_of(b, 0) + loss_of(b, 1)).backward()                         # summed loss, one backward
    rel = float((b.flat_gradient() - want).norm() / want.norm())
    assert rel < 2e-3, rel
    c = fresh()
    l0 = loss_of(c, 0)
    with torch.no_grad():
        c(pts[1], emb[1], idx)                                         # inference in between (own workspace)
    l0.backward(); loss_of(c, 1).backward()
    rel = float((c.flat_gradient() - want).norm() / want.norm())
    assert rel < 2e-3, rel


def test_get_surface_and_icp_regression():
    from autoposeestimation_b200.pc_reconstruction.open3d_utils import PointCloud, get_surface, icp_regression, icp_regression_batch
    fr = synth.render_ellipsoid_frame(6)
    surf = get_surface(fr['label'], fr['depth'].astype(np.float64), fr['intr'], fr['robot2cam'], 20, 5, 20, voxel_size=2,
                       outlier_filters=False)
    pts, _ = og.surface_backproject(fr['label'], fr['depth'].astype(np.float64), fr['intr'], fr['robot2cam'])
    want = oicp.voxel_down_sample(pts, 2.0)
    assert len(surf) == len(want) and np.allclose(surf.numpy(), want, atol=1e-9)
    # the reference's full chain (open3d_utils.py:198-211): voxel grid -> radius outliers -> statistical outliers
    full = get_surface(fr['label'], fr['depth'].astype(np.float64), fr['intr'], fr['robot2cam'], 20, 5, 20, voxel_size=2)
    want_full = oicp.get_surface_filters(pts, 20, 5, 20, 2.0)
    assert len(full) == len(want_full) and np.allclose(full.numpy(), want_full, atol=1e-9)
    tgt_d, src_d, T = icp_regression(PointCloud(fr['model']), surf, voxel_size=2, threshold=10, icp_point2plane=False)
    t_ref, s_ref, T_ref = oicp.icp_regression(fr['model'], surf.numpy(), voxel_size=2, threshold=10)
    assert len(tgt_d) == len(t_ref) and len(src_d) == len(s_ref)
    assert np.abs(T - T_ref).max() < 1e-5
    with pytest.raises(NotImplementedError):
        icp_regression(PointCloud(fr['model']), surf, global_regression=True, icp_point2plane=False)
    with pytest.raises(NotImplementedError):                      # the signature default icp_point2plane=True is never silently skipped
        icp_regression(PointCloud(fr['model']), surf, voxel_size=2, threshold=10)
    with pytest.raises(ValueError):
        get_surface(fr['label'], fr['depth'] + 0.5, fr['intr'], fr['robot2cam'], 20, 5, 20, voxel_size=2)
    with pytest.raises(ValueError):
        get_surface(fr['label'], fr['depth'].astype(np.float64), fr['intr'], fr['robot2cam'])         # filter parameters are required
    # merge step of create_pointcloud.py:307-312
    merged = PointCloud(torch.cat([src_d.transform(T).points, tgt_d.points])).voxel_down_sample(2)
    assert 0 < len(merged) <= len(src_d) + len(tgt_d)
    td, sd, Ts, info = icp_regression_batch([PointCloud(fr['model'])] * 3, [surf] * 3, voxel_size=2, threshold=10)
    assert np.abs(Ts - T_ref).max() < 1e-5 and info.shape == (3, 4)


def test_predict_poses_frame_block():
    """Option-6 geometry block for two objects of one frame vs the per-object reference flow (oracle)."""
    from autoposeestimation_b200.pipeline.utils import predict_poses, get_bbox, choose_points
    rng = np.random.RandomState(12)
    H, W, N, nobj = 480, 640, 1000, 3
    est, ref = _modules(9, N, nobj)
    enc_w = torch.randn(32, 3, device='cuda') * 0.5                      # stand-in encoder: per-pixel linear map 3 -> 32

    class _Enc(torch.nn.Module):
        def forward(self, crops):
            return torch.einsum('oc,bchw->bohw', enc_w, crops)
    est.cnn = _Enc()
    image = torch.randn(3, H, W, device='cuda')
    depth = rng.randint(400, 900, size=(H, W)).astype(np.uint16); depth[rng.rand(H, W) < 0.05] = 0
    masks = []
    for (r0, c0, h, w) in ((100, 150, 70, 90), (300, 400, 50, 130)):
        m = np.zeros((H, W), np.uint8); m[r0:r0 + h, c0:c0 + w] = 255; masks.append(m)
    masks.append(np.zeros((H, W), np.uint8))                              # empty mask -> skipped
    meta = {'intr': dict(ppx=320.0, ppy=240.0, fx=615.0, fy=615.0), 'depth_scale': 0.001}
    seed_state = np.random.RandomState(77)
    out = predict_poses(image, depth, meta, masks, [0, 2, 1], est, ref, num_points=N, refine_mode='live', rng=seed_state)
    assert sorted(out) == [0, 1]
    sd_e = synth.to_torch(synth.posenet_state_dict(9, nobj)); sd_r = synth.to_torch(synth.refiner_state_dict(1009, nobj))
    seed_state = np.random.RandomState(77)
    for i, cls in ((0, 0), (1, 2)):
        ml = masks[i] == 255
        bbox = get_bbox(ml)
        ch = choose_points(ml & (depth != 0), bbox, N, seed_state)
        cloud = og.backproject_choose(depth, bbox, ch, 320.0, 240.0, 615.0, 615.0, 0.001)
        crop = image[:, bbox[0]:bbox[1], bbox[2]:bbox[3]][None]
        out_img = est.cnn(crop).cpu()
        with torch.no_grad():
            res = odf.live_prediction(sd_e, sd_r, out_img, torch.from_numpy(cloud)[None], torch.from_numpy(ch.astype(np.int64))[None, None],
                                      torch.tensor([[cls]]), nobj, refine_calls=2)
        assert pm.rotation_angle_between(out[i]['rotation'], res['q']) < 1e-3
        assert np.abs(out[i]['position'] - res['t']).max() < 1e-4
    # device-side sampling: same flow, mask -> bbox -> choose -> cloud in one kernel, hashed subset
    out_d = predict_poses(image, depth, meta, masks, [0, 2, 1], est, ref, num_points=N, refine_mode='live', device_sampling=True, seed=5)
    assert sorted(out_d) == [0, 1]
    for i, cls in ((0, 0), (1, 2)):
        ml = masks[i] == 255
        bbox = og.get_bbox(ml)
        ch = og.choose_hashed(og.choose_candidates(ml, depth, bbox), N, (i * 0x9E3779B1 + 5) & 0xffffffff)
        cloud = og.backproject_choose(depth, bbox, ch, 320.0, 240.0, 615.0, 615.0, 0.001)
        out_img = est.cnn(image[:, bbox[0]:bbox[1], bbox[2]:bbox[3]][None]).cpu()
        with torch.no_grad():
            res = odf.live_prediction(sd_e, sd_r, out_img, torch.from_numpy(cloud)[None], torch.from_numpy(ch.astype(np.int64))[None, None],
                                      torch.tensor([[cls]]), nobj, refine_calls=2)
        assert pm.rotation_angle_between(out_d[i]['rotation'], res['q']) < 1e-3
        assert np.abs(out_d[i]['position'] - res['t']).max() < 1e-4
