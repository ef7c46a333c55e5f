"""CPU: the C-ABI library loads and exports every symbol include/ape_b200.h declares; host-side logic of the
drop-in (bbox, choose, state_dict contract, encoder, torch training path) against the oracle / golden vectors."""
import json
import os
import re
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, 'include', 'ape_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(ape_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from autoposeestimation_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 18
    exported = subprocess.run(['nm', '-D', '--defined-only', _lib.LIB_PATH], capture_output=True, text=True).stdout
    for s in syms:
        assert re.search(r'\bT %s\b' % s, exported), 'not exported: ' + s
        assert hasattr(lib, s)
    assert set(syms) == set(_lib.SIGNATURES), set(syms) ^ set(_lib.SIGNATURES)     # ctypes table == header
    assert lib.ape_version() >= 100


def test_no_cuda_means_loud_failure_not_fallback():
    from autoposeestimation_b200 import _lib, ops
    if torch.cuda.is_available():
        pytest.skip('needs a box without a GPU')
    with pytest.raises(_lib.ApeError):
        ops.knn(torch.zeros((1, 3, 8)), torch.zeros((1, 3, 4)), 1)
    from autoposeestimation_b200 import synthetic as synth
    with pytest.raises(_lib.ApeError):
        ops.NetHandle(ops.NET_REFINER, synth.refiner_state_dict(0, 2), 2, 1, 128)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'autoposeestimation_b200')
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', txt, flags=re.M), os.path.join(dp, f)


def test_get_bbox_and_choose_match_oracle():
    from autoposeestimation_b200.pipeline import utils as pu
    from oracle import geometry as og
    rng = np.random.RandomState(0)
    for _ in range(400):
        m = np.zeros((480, 640), bool)
        r0, c0 = rng.randint(0, 479), rng.randint(0, 639)
        m[r0:r0 + rng.randint(1, 481 - r0), c0:c0 + rng.randint(1, 641 - c0)] = True
        m &= rng.rand(480, 640) < 0.7
        if not m.any():
            continue
        assert pu.get_bbox(m) == og.get_bbox(m)
    m = np.zeros((480, 640), bool); m[100:140, 200:260] = True
    bbox = pu.get_bbox(m)
    depth = rng.randint(0, 3, size=(480, 640))
    for n in (10, 1000, 5000):
        a = pu.choose_points(m & (depth != 0), bbox, n, np.random.RandomState(5))
        cand = og.choose_candidates(m, depth, bbox)
        keep = og.make_keep(len(cand), n, np.random.RandomState(5)) if len(cand) > n else None
        assert np.array_equal(a, og.choose_fixed(cand, n, keep))
    assert pu.choose_points(np.zeros((480, 640), bool), (0, 40, 0, 40), 10) is None      # empty -> skip (:530)


def test_state_dict_contract_matches_reference(golden_dir):
    from autoposeestimation_b200.densefusion import network
    shapes = json.load(open(os.path.join(golden_dir, 'state_dict_shapes.json')))
    for name, mod in (('posenet', network.PoseNet(500, 5)), ('refiner', network.PoseRefineNet(500, 5))):
        mine = {k: list(v.shape) for k, v in mod.state_dict().items()}
        assert mine == shapes[name], set(mine) ^ set(shapes[name])


def test_colour_encoder_matches_reference(golden_dir):
    from autoposeestimation_b200 import synthetic as synth
    from autoposeestimation_b200.densefusion import network
    shapes = json.load(open(os.path.join(golden_dir, 'state_dict_shapes.json')))['posenet']
    g = np.load(os.path.join(golden_dir, 'encoder.npz'))
    enc = network.ModifiedResnet().eval()
    enc.load_state_dict(synth.to_torch(synth.encoder_state_dict(77, {k[4:]: v for k, v in shapes.items() if k.startswith('cnn.')})))
    torch.set_num_threads(4)
    with torch.no_grad():
        out = enc(torch.from_numpy(g['img'])).numpy()
    assert out.shape == (1, 32, 40, 56)
    assert np.allclose(out[:, :, ::4, ::4], g['out_sub4'], atol=2e-4, rtol=1e-4)


@pytest.mark.parametrize('case', [0, 1])
def test_training_path_matches_reference(golden_dir, case):
    """The torch (autograd) path of the drop-in modules reproduces the reference's outputs and is differentiable."""
    from autoposeestimation_b200 import synthetic as synth
    from autoposeestimation_b200.densefusion import network
    g = np.load(os.path.join(golden_dir, 'densefusion_case%d.npz' % case))
    seed, npts, nobj = int(g['seed']), int(g['npts']), int(g['nobj'])
    hw = tuple(int(v) for v in g['hw'])
    est = network.PoseNet(npts, nobj); ref = network.PoseRefineNet(npts, nobj)
    est.load_state_dict(synth.to_torch(synth.posenet_state_dict(seed, nobj)), strict=False)
    ref.load_state_dict(synth.to_torch(synth.refiner_state_dict(seed + 1000, nobj)), strict=True)
    est.train(); ref.train()
    out_img, cloud, choose, idx = (torch.from_numpy(a) for a in synth.posenet_inputs(seed, npts, hw, nobj))
    torch.set_num_threads(4)
    r, t, c, emb = est.forward_geometry(out_img, cloud, choose, idx)
    assert np.allclose(r.detach().numpy(), g['r'], atol=1e-5) and np.allclose(t.detach().numpy(), g['t'], atol=1e-5)
    assert np.allclose(c.detach().numpy(), g['c'], atol=1e-6)
    r2, t2 = ref(torch.from_numpy(g['new_points']), emb, idx)
    assert np.allclose(r2.detach().numpy(), g['r2'], atol=1e-5) and np.allclose(t2.detach().numpy(), g['t2'], atol=1e-5)
    (r2.sum() + t2.sum()).backward()
    assert ref.conv1_r.weight.grad is not None and float(ref.feat.conv5.weight.grad.abs().sum()) > 0


def test_loss_dropins_nonsymmetric_match_reference(golden_dir):
    """Loss / Loss_refine torch glue (non-symmetric branch runs without the kNN kernel) vs the reference's outputs."""
    from autoposeestimation_b200.densefusion.loss import Loss
    from autoposeestimation_b200.densefusion.loss_refiner import Loss_refine
    g = np.load(os.path.join(golden_dir, 'losses.npz'))
    T = lambda k: torch.from_numpy(g[k])
    dis, npn, ntg, pred = Loss_refine(120, [])(T('pr1'), T('pt1'), T('target'), T('model'), torch.LongTensor([[0]]), T('points'))
    assert np.allclose(dis.numpy(), g['lr_dis_nosym'], atol=1e-7) and np.allclose(npn.numpy(), g['lr_newp_nosym'], atol=1e-6)
    assert np.allclose(ntg.numpy(), g['lr_newt_nosym'], atol=1e-6) and np.allclose(pred.numpy(), g['lr_pred_nosym'], atol=1e-6)
    for tag, sym, refine in (('nosym', [], False), ('symrefine', [0], True)):
        lo, d, npn, ntg, _ = Loss(120, sym)(T('pr_n'), T('pt_n'), T('pc_n'), T('target'), T('model'), torch.LongTensor([[0]]),
                                            T('points'), 0.015, refine)
        assert np.allclose(lo.numpy(), g['l_loss_' + tag], atol=1e-6) and np.allclose(d.numpy(), g['l_dis_' + tag], atol=1e-6)
        assert np.allclose(npn.numpy(), g['l_newp_' + tag], atol=1e-6) and np.allclose(ntg.numpy(), g['l_newt_' + tag], atol=1e-6)


def test_host_quaternion_helpers(golden_dir):
    from autoposeestimation_b200.densefusion import transformations as tf
    g = np.load(os.path.join(golden_dir, 'pose_math.npz'))
    for q, M in zip(g['quats'], g['quat_mats']):
        assert np.allclose(tf.quaternion_matrix(q), M, atol=1e-15)
    for R, q in zip(g['rots'], g['rot_quats']):
        assert np.allclose(tf.quaternion_from_matrix(R, True), q, atol=1e-15)
        q2 = tf.quaternion_from_matrix(R, False)
        assert min(np.abs(q2 - q).max(), np.abs(q2 + q).max()) < 1e-8
