"""Generate tests/golden/*.npz by running the REFERENCE's own Python modules
(imported unmodified from /root/reference) on seeded synthetic inputs.

Run here (dev container) only:  python -m oracle.gen_golden
The fixtures store seeds + reference OUTPUTS; inputs/weights are regenerated from
the seeds by oracle/synth.py.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Reference modules exercised (no source is copied):
  DenseFusion/lib/network.py        PoseNet (cnn replaced by nn.Identity so that `img` is the
                                    encoder output), PoseRefineNet
  DenseFusion/tools/utils.py        my_estimator_prediction, my_refined_prediction, get_new_points
  DenseFusion/lib/transformations.py quaternion_matrix, quaternion_from_matrix
  DenseFusion/lib/loss.py, loss_refiner.py   Loss, Loss_refine (KNearestNeighbor rebound to the
                                    reference's own knn_cpu.cpp, because the legacy autograd
                                    Function cannot be called on torch 2.x and forces .cuda())
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

REF = os.environ.get('APE_REFERENCE', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')


def _import_reference():
    sys.path.insert(0, REF)
    warnings.filterwarnings('ignore')
    from oracle import clib

    class _KNN:                       # plain callable with the reference's call syntax
        def __init__(self, k):
            self.k = k

        def __call__(self, ref, query):
            out = clib.ref_knn_cpu(ref.detach().numpy(), query.detach().numpy(), self.k)
            return torch.from_numpy(out)

    stub = types.ModuleType('DenseFusion.lib.knn.knn_pytorch')
    stub.knn = lambda *a: 1
    sys.modules['DenseFusion.lib.knn.knn_pytorch'] = stub
    import DenseFusion.lib.network as network
    import DenseFusion.tools.utils as tools
    import DenseFusion.lib.transformations as tf
    import DenseFusion.lib.loss as loss
    import DenseFusion.lib.loss_refiner as loss_refiner
    loss.KNearestNeighbor = _KNN
    loss_refiner.KNearestNeighbor = _KNN
    return network, tools, tf, loss, loss_refiner


def main():
    from oracle import synth
    network, tools, tf, loss_m, lossr_m = _import_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(4)

    # ------------------------------------------------------------------ pose math
    rng = np.random.RandomState(100)
    quats = rng.standard_normal((32, 4))
    quats[0] = [1, 0, 0, 0]; quats[1] = [0, 1, 0, 0]; quats[2] = [0, 0, 0, 1e-3]
    qm = np.stack([tf.quaternion_matrix(q) for q in quats])
    rots = np.stack([tf.quaternion_matrix(q / np.linalg.norm(q)) for q in rng.standard_normal((32, 4))])
    # force every branch of quaternion_from_matrix: near-180-degree rotations about each axis
    for i, ax in enumerate(([1, 0, 0], [0, 1, 0], [0, 0, 1])):
        rots[i] = tf.rotation_matrix(np.pi - 0.01 * (i + 1), ax)
    qfm = np.stack([tf.quaternion_from_matrix(np.array(R), True) for R in rots])
    ref_in = dict(r2=rng.standard_normal((16, 4)).astype(np.float32), t2=(rng.standard_normal((16, 3)) * 0.05).astype(np.float32),
                  my_r=None, my_t=(rng.standard_normal((16, 3)) * 0.3).astype(np.float32))
    mr = rng.standard_normal((16, 4)); mr /= np.linalg.norm(mr, axis=1, keepdims=True)
    ref_in['my_r'] = mr.astype(np.float32)
    ref_q, ref_t = [], []
    for i in range(16):
        _, q, t = tools.my_refined_prediction(torch.from_numpy(ref_in['r2'][i:i + 1]), torch.from_numpy(ref_in['t2'][i:i + 1]),
                                              ref_in['my_r'][i], ref_in['my_t'][i])
        ref_q.append(q); ref_t.append(t)
    n = 200
    pr = rng.standard_normal((1, n, 4)).astype(np.float32); pt = (rng.standard_normal((1, n, 3)) * 0.1).astype(np.float32)
    pc = rng.uniform(0.05, 0.95, size=(1, n, 1)).astype(np.float32); pts = rng.standard_normal((1, n, 3)).astype(np.float32)
    newp = tools.get_new_points(*(torch.from_numpy(a) for a in (pr, pt, pc, pts))).numpy()
    _, er, et = tools.my_estimator_prediction(torch.from_numpy(pr), torch.from_numpy(pt), torch.from_numpy(pc), n, 1, torch.from_numpy(pts))
    np.savez_compressed(os.path.join(OUT, 'pose_math.npz'), quats=quats, quat_mats=qm, rots=rots, rot_quats=qfm,
                        r2=ref_in['r2'], t2=ref_in['t2'], my_r=ref_in['my_r'], my_t=ref_in['my_t'],
                        refined_q=np.array(ref_q), refined_t=np.array(ref_t),
                        pr=pr, pt=pt, pc=pc, pts=pts, new_points=newp, est_r=er, est_t=et)

    # ------------------------------------------------------------------ networks
    cases = []
    for case, (seed, npts, hw, nobj) in enumerate([(11, 64, (40, 40), 3), (12, 500, (120, 160), 5), (13, 1000, (80, 120), 5)]):
        est = network.PoseNet(npts, nobj); est.cnn = torch.nn.Identity(); est.eval()
        refn = network.PoseRefineNet(npts, nobj); refn.eval()
        sd_e = synth.to_torch(synth.posenet_state_dict(seed, nobj))
        sd_r = synth.to_torch(synth.refiner_state_dict(seed + 1000, nobj))
        est.load_state_dict(sd_e, strict=True); refn.load_state_dict(sd_r, strict=True)
        out_img, cloud, choose, idx = (torch.from_numpy(a) for a in synth.posenet_inputs(seed, npts, hw, nobj))
        with torch.no_grad():
            r, t, c, emb = est(out_img, cloud, choose, idx)
            newp = tools.get_new_points(r, t, c, cloud)
            _, my_r, my_t = tools.my_estimator_prediction(r, t, c, npts, 1, cloud)
            r2, t2 = refn(newp, emb, idx)
            _, fq, ft = tools.my_refined_prediction(r2, t2, my_r, my_t)
        cases.append(dict(seed=seed, npts=npts, h=hw[0], w=hw[1], nobj=nobj))
        np.savez_compressed(os.path.join(OUT, 'densefusion_case%d.npz' % case), seed=seed, npts=npts, hw=np.array(hw), nobj=nobj,
                            r=r.numpy(), t=t.numpy(), c=c.numpy(), emb_sum=emb.numpy().sum(axis=1), new_points=newp.numpy(),
                            my_r=my_r, my_t=my_t, r2=r2.numpy(), t2=t2.numpy(), final_q=fq, final_t=ft)

    # ------------------------------------------------------------------ state_dict contract + colour encoder
    import json
    est = network.PoseNet(500, 5); refn = network.PoseRefineNet(500, 5)
    shapes = dict(posenet={k: list(v.shape) for k, v in est.state_dict().items()},
                  refiner={k: list(v.shape) for k, v in refn.state_dict().items()})
    with open(os.path.join(OUT, 'state_dict_shapes.json'), 'w') as f:
        json.dump(shapes, f, indent=0, sort_keys=True)
    enc_sd = synth.encoder_state_dict(77, {k[len('cnn.'):]: v for k, v in shapes['posenet'].items() if k.startswith('cnn.')})
    est.cnn.load_state_dict(synth.to_torch(enc_sd), strict=True)
    est.cnn.eval()
    rng = np.random.RandomState(78)
    img = rng.standard_normal((1, 3, 40, 56)).astype(np.float32)
    with torch.no_grad():
        enc_out = est.cnn(torch.from_numpy(img)).numpy()
    np.savez_compressed(os.path.join(OUT, 'encoder.npz'), img=img, out_sub4=enc_out[:, :, ::4, ::4])   # small fixture

    # ------------------------------------------------------------------ losses (ADD / ADD-S)
    rng = np.random.RandomState(300)
    m = 120; npt = 60
    model = rng.uniform(-0.1, 0.1, size=(1, m, 3)).astype(np.float32)
    q_gt = rng.standard_normal(4); q_gt /= np.linalg.norm(q_gt)
    R_gt = tf.quaternion_matrix(q_gt)[:3, :3]
    t_gt = np.array([0.05, -0.02, 0.6])
    target = (model[0] @ R_gt.T + t_gt).astype(np.float32)[None]
    points = (target[0][rng.choice(m, npt)] + rng.standard_normal((npt, 3)) * 1e-3).astype(np.float32)[None]
    pr1 = (q_gt + rng.standard_normal(4) * 0.02).astype(np.float32)[None]
    pt1 = (t_gt + rng.standard_normal(3) * 0.005).astype(np.float32)[None]
    res = {}
    for tag, sym in (('sym', [0]), ('nosym', [])):
        L = lossr_m.Loss_refine(m, sym)
        dis, npn, ntg, pred = L(torch.from_numpy(pr1), torch.from_numpy(pt1), torch.from_numpy(target), torch.from_numpy(model),
                                torch.LongTensor([[0]]), torch.from_numpy(points))
        res.update({'lr_dis_' + tag: dis.numpy(), 'lr_newp_' + tag: npn.numpy(), 'lr_newt_' + tag: ntg.numpy(), 'lr_pred_' + tag: pred.numpy()})
    pr_n = (q_gt[None] + rng.standard_normal((npt, 4)) * 0.05).astype(np.float32)[None]
    pt_n = ((t_gt - points[0]) + rng.standard_normal((npt, 3)) * 0.005).astype(np.float32)[None]
    pc_n = rng.uniform(0.1, 0.9, size=(1, npt, 1)).astype(np.float32)
    for tag, sym, refine in (('sym', [0], False), ('nosym', [], False), ('symrefine', [0], True)):
        L = loss_m.Loss(m, sym)
        lo, dis, npn, ntg, pred = L(torch.from_numpy(pr_n), torch.from_numpy(pt_n), torch.from_numpy(pc_n), torch.from_numpy(target),
                                    torch.from_numpy(model), torch.LongTensor([[0]]), torch.from_numpy(points), 0.015, refine)
        res.update({'l_loss_' + tag: lo.numpy(), 'l_dis_' + tag: dis.numpy(), 'l_newp_' + tag: npn.numpy(), 'l_newt_' + tag: ntg.numpy()})
    np.savez_compressed(os.path.join(OUT, 'losses.npz'), model=model, target=target, points=points, pr1=pr1, pt1=pt1,
                        pr_n=pr_n, pt_n=pt_n, pc_n=pc_n, **res)

    # ------------------------------------------------------------------ kNN (reference knn_cpu.cpp, unmodified)
    from oracle import clib
    rng = np.random.RandomState(400)
    ref = rng.uniform(-0.1, 0.1, size=(2, 3, 300)).astype(np.float32)
    qry = rng.uniform(-0.1, 0.1, size=(2, 3, 77)).astype(np.float32)
    ref[:, :, 17] = ref[:, :, 5]                       # an exact duplicate -> exact-tie rule
    qry[:, :, 3] = ref[:, :, 5]
    refd = rng.rand(1, 16, 90).astype(np.float32); qryd = rng.rand(1, 16, 33).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, 'knn.npz'), ref=ref, qry=qry, idx_k1=clib.ref_knn_cpu(ref, qry, 1),
                        idx_k4=clib.ref_knn_cpu(ref, qry, 4), refd=refd, qryd=qryd, idxd_k2=clib.ref_knn_cpu(refd, qryd, 2))
    print('golden fixtures written to', os.path.abspath(OUT))
    for f in sorted(os.listdir(OUT)):
        print('  %-28s %8d B' % (f, os.path.getsize(os.path.join(OUT, f))))


def loss_grads():
    """tests/golden/loss_grads.npz: the REFERENCE's own `Loss` (lib/loss.py:12-73) differentiated by autograd -- loss value
    and d loss / d (pred_r, pred_t, pred_c) for a symmetric and a non-symmetric class (inputs: tests/golden/losses.npz)."""
    network, tools, tf, loss_m, lossr_m = _import_reference()
    g = np.load(os.path.join(OUT, 'losses.npz'))
    res = {}
    for tag, sym in (('sym', [0]), ('nosym', [])):
        pr = torch.from_numpy(g['pr_n']).requires_grad_(True); pt = torch.from_numpy(g['pt_n']).requires_grad_(True)
        pc = torch.from_numpy(g['pc_n']).requires_grad_(True)
        L = loss_m.Loss(g['model'].shape[1], sym)
        lo, dis, npn, ntg, pred = L(pr, pt, pc, torch.from_numpy(g['target']), torch.from_numpy(g['model']), torch.LongTensor([[0]]),
                                    torch.from_numpy(g['points']), 0.015, False)
        lo.backward()
        res.update({'loss_' + tag: lo.detach().numpy(), 'd_r_' + tag: pr.grad.numpy(), 'd_t_' + tag: pt.grad.numpy(),
                    'd_c_' + tag: pc.grad.numpy(), 'pred_' + tag: pred.detach().numpy()})
    np.savez_compressed(os.path.join(OUT, 'loss_grads.npz'), **res)
    print('loss_grads.npz written', {k: v.shape for k, v in res.items()})


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'grads':
        loss_grads()
    else:
        main()
