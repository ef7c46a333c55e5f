"""GPU parity: back-projection kernels vs the oracle (bit-exact indices / fp32, fp64 within 1e-9 mm)."""
import numpy as np
import pytest
import torch

from oracle import geometry as og, synth

pytestmark = pytest.mark.gpu


def _dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def _u16(a):
    return torch.from_numpy(a.astype(np.uint16).view(np.int16).copy()).cuda()   # same bits; kernel reads uint16


@pytest.mark.parametrize('seed,num_points', [(0, 500), (1, 1000), (2, 20000)])
def test_backproject_choose_bit_exact(seed, num_points):
    from autoposeestimation_b200 import ops
    fr = synth.render_ellipsoid_frame(seed)
    depth, label = fr['depth'], fr['label']
    bbox = og.get_bbox(label == 255)
    cand = og.choose_candidates(label == 255, depth, bbox)
    rng = np.random.RandomState(seed)
    keep = og.make_keep(len(cand), num_points, rng) if len(cand) > num_points else None
    choose = og.choose_fixed(cand, num_points, keep)            # num_points=20000 exercises the 'wrap' padding
    intr = fr['intr']
    ref = og.backproject_choose(depth, bbox, choose, intr['ppx'], intr['ppy'], intr['fx'], intr['fy'], synth.DEPTH_SCALE)
    cam = np.array([[intr['ppx'], intr['ppy'], intr['fx'], intr['fy'], synth.DEPTH_SCALE]], np.float32)
    out = ops.backproject_choose(_u16(depth[None]), _dev(np.array([bbox], np.int32)), _dev(choose[None].astype(np.int64)), _dev(cam))
    got = out.cpu().numpy()[0]
    assert got.dtype == np.float32 and np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_backproject_choose_batch_and_border_bbox():
    from autoposeestimation_b200 import ops
    rng = np.random.RandomState(3)
    F, H, W, N = 3, 480, 640, 300
    depth = rng.randint(0, 4000, size=(F, H, W)).astype(np.uint16)
    bboxes, chooses, frames, refs = [], [], [], []
    for b in range(5):
        m = np.zeros((H, W), bool)
        r0, c0 = [(0, 0), (440, 600), (200, 300), (0, 620), (470, 0)][b]
        m[r0:r0 + 10 + 7 * b, c0:c0 + 15 + 3 * b] = True           # touches the image borders -> clamped bbox
        bbox = og.get_bbox(m)
        f = b % F
        cand = og.choose_candidates(m, depth[f], bbox)
        ch = og.choose_fixed(cand, N, og.make_keep(len(cand), N, rng) if len(cand) > N else None)
        bboxes.append(bbox); chooses.append(ch); frames.append(f)
        refs.append(og.backproject_choose(depth[f], bbox, ch, 321.5, 238.25, 614.7, 615.3, 0.001))
    cam = np.tile(np.array([[321.5, 238.25, 614.7, 615.3, 0.001]], np.float32), (5, 1))
    out = ops.backproject_choose(_u16(depth), _dev(np.array(bboxes, np.int32)), _dev(np.array(chooses, np.int64)), _dev(cam),
                                 frame_of=_dev(np.array(frames, np.int32))).cpu().numpy()
    for b in range(5):
        assert np.array_equal(out[b].view(np.uint32), refs[b].view(np.uint32))


def test_get_bbox_oracle_rules():
    # dataset.py:342-380 rules (host logic of the drop-in is the same function; see test_host_logic)
    m = np.zeros((480, 640), bool); m[100:139, 50:131] = True      # 39 rows -> 40, 81 cols -> 120
    assert og.get_bbox(m) == (99, 139, 30, 150)
    m = np.zeros((480, 640), bool); m[0:5, 0:5] = True
    assert og.get_bbox(m) == (0, 40, 0, 40)
    m = np.zeros((480, 640), bool); m[470:480, 630:640] = True
    assert og.get_bbox(m) == (440, 480, 600, 640)


@pytest.mark.parametrize('seed', [0, 5])
def test_surface_backproject_vs_oracle(seed):
    from autoposeestimation_b200 import ops
    fr = synth.render_ellipsoid_frame(seed)
    ref_pts, ref_pix = og.surface_backproject(fr['label'], fr['depth'].astype(np.float64), fr['intr'], fr['robot2cam'])
    intr = fr['intr']
    cam = np.array([[intr['ppx'], intr['ppy'], intr['fx'], intr['fy']]], np.float64)
    pts, pix, cnt = ops.surface_backproject(_dev(fr['label'][None]), _u16(fr['depth'][None]), _dev(cam), _dev(fr['robot2cam'][None]),
                                            capacity=len(ref_pix) + 100)
    n = int(cnt.cpu()[0])
    assert n == len(ref_pix)
    assert np.array_equal(pix.cpu().numpy()[0, :n], ref_pix.astype(np.int32))          # indices bit-exact, row-major order
    assert np.allclose(pts.cpu().numpy()[0, :n], ref_pts, rtol=0, atol=1e-9)


def test_surface_backproject_literal_loop_small():
    """The reference's literal per-pixel loop (open3d_utils.py:175-192) on a small mask."""
    from autoposeestimation_b200 import ops
    rng = np.random.RandomState(9)
    H, W = 480, 640
    label = np.zeros((H, W), np.uint8); label[200:215, 310:330] = 255
    depth = rng.randint(0, 3, size=(H, W)).astype(np.uint16) * rng.randint(400, 900, size=(H, W)).astype(np.uint16)
    T = synth.hand_eye()
    ref_pts, ref_pix = og.surface_backproject_literal(label, depth.astype(np.float64), synth.INTR, T)
    cam = np.array([[synth.INTR['ppx'], synth.INTR['ppy'], synth.INTR['fx'], synth.INTR['fy']]], np.float64)
    pts, pix, cnt = ops.surface_backproject(_dev(label[None]), _u16(depth[None]), _dev(cam), _dev(T[None]), capacity=400)
    n = int(cnt.cpu()[0])
    assert n == len(ref_pix) and np.array_equal(pix.cpu().numpy()[0, :n], ref_pix.astype(np.int32))
    assert np.allclose(pts.cpu().numpy()[0, :n], ref_pts, rtol=0, atol=1e-9)


def test_surface_backproject_views_labels_capacity_empty():
    from autoposeestimation_b200 import ops
    rng = np.random.RandomState(11)
    F, H, W = 2, 480, 640
    label = np.zeros((F, H, W), np.uint8)
    label[0, 10:60, 20:90] = 1; label[0, 300:330, 500:640] = 2; label[1, 470:480, 0:640] = 3   # last rows / full width
    depth = rng.randint(0, 2000, size=(F, H, W)).astype(np.uint16)
    depth[rng.rand(F, H, W) < 0.1] = 0
    T = np.stack([synth.hand_eye(), np.identity(4)])
    views = [(0, 1), (0, 2), (1, 3), (1, 0), (0, 7)]           # (frame, label value); 0 = any non-zero; 7 = empty
    cam = np.tile(np.array([[320.0, 240.0, 615.0, 615.0]]), (len(views), 1))
    r2c = np.stack([T[f] for f, _ in views])
    cap = 3000
    pts, pix, cnt = ops.surface_backproject(_dev(label), _u16(depth), _dev(cam), _dev(r2c), capacity=cap,
                                            frame_of=_dev(np.array([f for f, _ in views], np.int32)),
                                            label_value=_dev(np.array([v for _, v in views], np.uint8)))
    cnt = cnt.cpu().numpy(); pix = pix.cpu().numpy(); pts = pts.cpu().numpy()
    for v, (f, val) in enumerate(views):
        lab = (label[f] == val) if val else (label[f] != 0)
        ref_pts, ref_pix = og.surface_backproject(lab.astype(np.uint8), depth[f].astype(np.float64),
                                                  dict(ppx=320.0, ppy=240.0, fx=615.0, fy=615.0), T[f])
        assert cnt[v] == len(ref_pix)                           # full count even when it exceeds capacity
        n = min(cap, len(ref_pix))
        assert np.array_equal(pix[v, :n], ref_pix[:n].astype(np.int32))
        assert np.allclose(pts[v, :n], ref_pts[:n], rtol=0, atol=1e-9)
    assert cnt[4] == 0 and cnt[1] > cap                         # empty view; overflowing view


def test_surface_backproject_full_batch_property():
    """BASELINE-size batch (64 frames): counts equal a torch reduction of the same predicate."""
    from autoposeestimation_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(0)
    F, H, W = 64, 480, 640
    label = (torch.rand((F, H, W), device='cuda', generator=g) < 0.04).to(torch.uint8) * 255
    depth = torch.randint(0, 1500, (F, H, W), device='cuda', generator=g, dtype=torch.int32)
    depth[depth < 30] = 0
    depth16 = depth.to(torch.int16)
    cam = torch.tensor([[320.0, 240.0, 615.0, 615.0]], dtype=torch.float64, device='cuda').repeat(F, 1)
    r2c = torch.eye(4, dtype=torch.float64, device='cuda').repeat(F, 1, 1)
    pts, pix, cnt = ops.surface_backproject(label, depth16, cam, r2c, capacity=16384)
    want = ((label != 0) & (depth != 0)).flatten(1).sum(1).to(torch.int32)
    assert torch.equal(cnt, want)
    n0 = int(cnt[0])
    p0 = pix[0, :n0].long()
    assert bool((p0[1:] > p0[:-1]).all())                       # strictly increasing = row-major order
    z = pts[0, :n0, 2]
    assert torch.equal(z, depth[0].flatten()[p0].double())      # identity extrinsic: z is the raw depth


@pytest.mark.parametrize('num_points', [500, 1000, 20000])
def test_mask_bbox_choose_on_device_vs_oracle(num_points):
    """label -> get_bbox -> choose -> back-projection in one kernel, bit-exact against the oracle with
    the same hash keys (subset branch for 500 / 1000 points, 'wrap' branch for 20000), several objects of one frame,
    a border-touching object, a label value other than 255, and an object without valid depth."""
    from autoposeestimation_b200 import ops
    rng = np.random.RandomState(7)
    H, W = 480, 640
    label = np.zeros((2, H, W), np.uint8)
    yy, xx = np.mgrid[0:H, 0:W]
    label[0][((yy - 200) / 50.0) ** 2 + ((xx - 300) / 70.0) ** 2 < 1] = 255         # ~11 k px
    label[0][((yy - 30) / 45.0) ** 2 + ((xx - 600) / 60.0) ** 2 < 1] = 7            # touches the top / right border, value 7
    label[1][100:140, 20:75] = 255                                                  # small box (wrap at 20000 and 1000? 2200 px)
    label[1][400:420, 400:430] = 9                                                  # will have zero depth everywhere
    depth = rng.randint(400, 900, size=(2, H, W)).astype(np.uint16)
    depth[rng.rand(2, H, W) < 0.05] = 0
    depth[1, 400:420, 400:430] = 0
    frame_of = np.array([0, 0, 1, 1], np.int32); values = np.array([255, 7, 255, 9], np.uint8)
    seeds = np.array([1, 123456789, 0xdeadbeef, 5], np.int64)
    cam = np.tile(np.array([[synth.INTR['ppx'], synth.INTR['ppy'], synth.INTR['fx'], synth.INTR['fy'], synth.DEPTH_SCALE]], np.float32), (4, 1))
    out = ops.mask_bbox_choose(_dev(label), _dev(depth.view(np.int16)), _dev(cam), num_points, frame_of=_dev(frame_of),
                               label_value=_dev(values), seeds=_dev(seeds))
    bbox = out['bbox'].cpu().numpy(); ncand = out['n_candidates'].cpu().numpy()
    choose = out['choose'].cpu().numpy(); cloud = out['cloud'].cpu().numpy()
    for b in range(4):
        f = frame_of[b]
        mask_label = label[f] == values[b]
        want_bbox = og.get_bbox(mask_label)
        cand = og.choose_candidates(mask_label, depth[f], want_bbox)
        assert ncand[b] == len(cand)
        if len(cand) == 0:
            continue                                                                # skipped object (:530-531)
        assert tuple(bbox[b]) == tuple(want_bbox)
        want = og.choose_hashed(cand, num_points, int(seeds[b]))
        assert np.array_equal(choose[b], want)                                      # bit-exact indices
        ref_cloud = og.backproject_choose(depth[f], want_bbox, want, cam[b, 0], cam[b, 1], cam[b, 2], cam[b, 3], cam[b, 4])
        assert np.array_equal(cloud[b].view(np.uint32), ref_cloud.view(np.uint32))  # bit-exact fp32 points
    assert ncand[3] == 0 and ncand[0] > 1000
    # the subset really is a subset without repeats, and different seeds give different subsets
    if num_points < ncand[0]:
        assert len(np.unique(choose[0])) == num_points
        out2 = ops.mask_bbox_choose(_dev(label), _dev(depth.view(np.int16)), _dev(cam), num_points, frame_of=_dev(frame_of),
                                    label_value=_dev(values), seeds=_dev(seeds + 1))
        assert not np.array_equal(out2['choose'].cpu().numpy()[0], choose[0])


# ---- against vectors produced by the reference's own code (oracle/gen_golden_geometry.py)
def test_backproject_choose_vs_reference_run(golden_dir):
    """DenseFusion/datasets/myDatasetAugmented/dataset.py:236-275 executed by the reference: fp32 cloud bit-exact for
    the reference's own `choose`, in one batched launch over the three frames."""
    import os
    from autoposeestimation_b200 import ops
    g = np.load(os.path.join(golden_dir, 'geometry_ref.npz'))
    ppx, ppy, fx, fy = (float(v) for v in g['intr'])
    F = len(g['seeds'])
    cam = np.tile(np.array([[ppx, ppy, fx, fy, float(g['depth_scale'])]], np.float32), (F, 1))
    out = ops.backproject_choose(_u16(g['depth']), _dev(g['bbox'].astype(np.int32)), _dev(g['choose'].astype(np.int64)), _dev(cam),
                                 frame_of=_dev(np.arange(F, dtype=np.int32))).cpu().numpy()
    assert np.array_equal(out.view(np.uint32), g['cloud'].view(np.uint32))


def test_surface_backproject_vs_reference_run(golden_dir):
    """pc_reconstruction/open3d_utils.py:171-192 executed by the reference: same number of points, same order, 1e-9 mm."""
    import os
    from autoposeestimation_b200 import ops
    g = np.load(os.path.join(golden_dir, 'geometry_ref.npz'))
    F = len(g['surf_n'])
    cam = np.tile(g['intr'][None].astype(np.float64), (F, 1))
    T = np.tile(g['robot2cam'][None], (F, 1, 1))
    pts, pix, cnt = ops.surface_backproject(_dev(g['label']), _u16(g['depth']), _dev(cam), _dev(T), capacity=int(g['surf_n'].max()) + 8)
    off = np.concatenate([[0], np.cumsum(g['surf_n'])])
    cnt = cnt.cpu().numpy(); pts = pts.cpu().numpy()
    for i in range(F):
        assert int(cnt[i]) == int(g['surf_n'][i])
        assert np.abs(pts[i, :cnt[i]] - g['surf_pts'][off[i]:off[i + 1]]).max() < 1e-9


def test_mask_bbox_choose_on_device_vs_reference_run(golden_dir):
    """Device-side mask -> bbox -> choose (csrc/choose.cu) on the reference-run frames: the bbox equals the reference's
    get_bbox; in the 'wrap' branch (frame 2: fewer candidates than num_pt) choose and cloud equal the reference's bit
    for bit; in the subset branch the chosen pixels are an ascending num_pt-subset of the reference's candidates."""
    import os
    from autoposeestimation_b200 import ops
    g = np.load(os.path.join(golden_dir, 'geometry_ref.npz'))
    ppx, ppy, fx, fy = (float(v) for v in g['intr'])
    F, N = len(g['seeds']), int(g['num_pt'])
    cam = np.tile(np.array([[ppx, ppy, fx, fy, float(g['depth_scale'])]], np.float32), (F, 1))
    res = ops.mask_bbox_choose(_dev(g['label']), _u16(g['depth']), _dev(cam), N, frame_of=_dev(np.arange(F, dtype=np.int32)),
                               seeds=_dev(np.array([9, 10, 11], np.int64)))
    bbox, choose, cloud = res['bbox'].cpu().numpy(), res['choose'].cpu().numpy(), res['cloud'].cpu().numpy()
    assert np.array_equal(bbox, g['bbox'].astype(bbox.dtype))
    assert np.array_equal(choose[2], g['choose'][2]) and np.array_equal(cloud[2].view(np.uint32), g['cloud'][2].view(np.uint32))
    for i in (0, 1):
        mask = (g['label'][i] == 255) & (g['depth'][i] != 0)
        r0, r1, c0, c1 = (int(v) for v in g['bbox'][i])
        cand = np.flatnonzero(mask[r0:r1, c0:c1].ravel())
        assert len(np.unique(choose[i])) == N and np.all(np.diff(choose[i]) > 0) and np.isin(choose[i], cand).all()


def test_surface_backproject_large_and_odd_frames():
    """Frames larger than 512 Ki pixels (the scan kernel keeps 16 x 32 sub-chunk counts in registers and loops over the
    rest), a frame whose size is not a multiple of the 8192-pixel chunk (ragged last chunk), several labels per frame and
    a capacity smaller than the number of valid pixels: indices bit-exact, points within 1e-9 of the oracle."""
    from autoposeestimation_b200 import ops
    rng = np.random.RandomState(11)
    for (H, W) in ((1200, 1600), (72, 1000), (16, 16)):
        label = np.zeros((2, H, W), np.uint8)
        label[0, H // 5:H // 2, W // 4:W // 2] = 255
        label[0, -3:, :] = 255                                        # last rows: the ragged tail of the frame
        label[0, 0, :7] = 255
        label[1, H // 3:, : W // 3] = 9
        depth = rng.randint(0, 3000, size=(2, H, W)).astype(np.uint16)
        depth[rng.rand(2, H, W) < 0.1] = 0
        intr = dict(ppx=W / 2 - 0.3, ppy=H / 2 + 0.2, fx=611.5, fy=612.25)
        r2c = synth.hand_eye()
        views = [(0, 255), (1, 9), (0, 9)]                             # the last one has no pixel at all
        cam = np.tile(np.array([[intr['ppx'], intr['ppy'], intr['fx'], intr['fy']]]), (len(views), 1))
        want = [og.surface_backproject(label[f] == v, depth[f].astype(np.float64), intr, r2c) for f, v in views]
        cap = max(len(w[1]) for w in want) + 5
        for capacity in (cap, max(1, len(want[0][1]) // 2)):
            pts, pix, cnt = ops.surface_backproject(_dev(label), _u16(depth), _dev(cam), _dev(np.tile(r2c[None], (len(views), 1, 1))),
                                                    capacity=capacity, frame_of=_dev(np.array([f for f, _ in views], np.int32)),
                                                    label_value=_dev(np.array([v for _, v in views], np.uint8)))
            pts = pts.cpu().numpy(); pix = pix.cpu().numpy(); cnt = cnt.cpu().numpy()
            for i, (wp, wi) in enumerate(want):
                assert cnt[i] == len(wi)                                # the count is the full number of valid pixels
                n = min(len(wi), capacity)
                assert np.array_equal(pix[i, :n], wi[:n].astype(np.int32))
                if n:
                    assert np.abs(pts[i, :n] - wp[:n]).max() < 1e-9


def test_surface_backproject_multi_label_packed_pipeline():
    """BASELINE config 4 building block: frames carrying the labels of 5 objects.  ONE pass per frame (multi-label mask
    kernel, packed ragged output) == the per-(frame, label) views of the single-label path, bit for bit, == the oracle's
    per-pixel loop; the packed offsets feed the voxel grid and (through src_count, without repacking) the ICP kernel, whose
    transforms equal those of the repacked call and the oracle's."""
    from autoposeestimation_b200 import ops, synthetic as psynth
    from oracle import icp as oicp
    scene = psynth.Scene(3)
    F, L = 4, 5
    T = scene.camera_poses(2, F)
    lab, dep = scene.render(T, seed=1, device='cuda')
    cam = torch.tensor([[psynth.INTR['ppx'], psynth.INTR['ppy'], psynth.INTR['fx'], psynth.INTR['fy']]], dtype=torch.float64, device='cuda').repeat(F, 1)
    r2c = torch.from_numpy(T).cuda()
    cap = 16384
    out = ops.surface_backproject_multi(lab, dep, cam, r2c, [1, 2, 3, 4, 5], total_capacity=F * L * 8192, want_pixels=True)
    frame_of = torch.arange(F, device='cuda', dtype=torch.int32).repeat_interleave(L)
    values = torch.arange(1, L + 1, device='cuda', dtype=torch.uint8).repeat(F)
    pts, pix, cnt = ops.surface_backproject(lab, dep, cam[frame_of.long()], r2c[frame_of.long()], capacity=cap, frame_of=frame_of, label_value=values)
    off = out['offsets'].cpu().numpy(); cn = cnt.cpu().numpy()
    assert np.array_equal(out['counts'].cpu().numpy(), cn) and np.array_equal(np.diff(off), cn) and off[0] == 0
    assert cn.min() > 500                                               # every object is visible in every frame
    for v in range(F * L):
        assert torch.equal(out['points'][off[v]:off[v + 1]], pts[v, :cn[v]]) and torch.equal(out['pixel_index'][off[v]:off[v + 1]], pix[v, :cn[v]])
    f, l = 2, 3
    want, wpix = og.surface_backproject(((lab[f] == l + 1).cpu().numpy() * 255).astype(np.uint8), dep[f].cpu().numpy().view(np.uint16).astype(np.float64),
                                        psynth.INTR, T[f])
    v = f * L + l
    assert np.array_equal(out['pixel_index'][off[v]:off[v + 1]].cpu().numpy(), wpix.astype(np.int32))
    assert np.allclose(out['points'][off[v]:off[v + 1]].cpu().numpy(), want, rtol=0, atol=1e-9)
    # voxel grid on the packed cloud, ICP straight from its gapped output
    vox, vc = ops.voxel_down_sample(out['points'], out['offsets'], 2.0)
    tgt = torch.from_numpy(np.concatenate([scene.models_pert[v % L] for v in range(F * L)])).cuda()
    to = torch.arange(0, F * L + 1, device='cuda', dtype=torch.int32) * 2000
    T1, info1 = ops.icp_p2p(vox, out['offsets'], tgt, to, 10.0, src_count=vc)
    vch = vc.cpu().numpy()
    packed = torch.cat([vox[off[v]:off[v] + vch[v]] for v in range(F * L)])
    po = np.zeros(F * L + 1, np.int32); po[1:] = np.cumsum(vch)
    T2, info2 = ops.icp_p2p(packed, torch.from_numpy(po).cuda(), tgt, to, 10.0)
    assert torch.equal(T1, T2) and torch.equal(info1, info2)
    src = oicp.voxel_down_sample(out['points'][off[v]:off[v + 1]].cpu().numpy(), 2.0)     # bit-exact on identical input points
    assert np.array_equal(vox[off[v]:off[v] + vch[v]].cpu().numpy(), src)
    T_ref = oicp.registration_icp_p2p(src, scene.models_pert[l], 10.0)
    assert np.abs(T1[v].cpu().numpy() - T_ref).max() < 1e-5
    # the registration brings the perturbed model back onto the observed surface: residual of the true model under T^-1
    assert float(info1[:, 0].min()) > 0.9                                # fitness
