"""Oracle (test infrastructure): DenseFusion PoseNet / PoseRefineNet geometry side,
losses and the refinement loops, on torch-CPU fp32.

Functional restatements driven by a reference-shaped ``state_dict``:
  * PoseNetFeat.forward            DenseFusion/lib/network.py:53-68
  * PoseNet.forward (after cnn)    DenseFusion/lib/network.py:98-132
  * PoseRefineNetFeat.forward      DenseFusion/lib/network.py:151-168
  * PoseRefineNet.forward          DenseFusion/lib/network.py:187-206
  * Loss_refine (ADD / ADD-S)      DenseFusion/lib/loss_refiner.py:12-64
  * Loss                           DenseFusion/lib/loss.py:12-73
  * live loop quirk                pipeline/utils.py:564-571
  * canonical refinement loop      DenseFusion/tools/eval_linemod.py:81-114
Pinned by golden vectors produced by importing the reference's own modules
(oracle/gen_golden.py -> tests/golden/densefusion_*.npz).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import pose_math


def _c1(sd, name, x, relu=True):
    w = sd[name + '.weight']
    if w.dim() == 2:
        w = w[:, :, None]
    y = F.conv1d(x, w, sd[name + '.bias'])
    return F.relu(y) if relu else y


def posenet_feat(sd, x, emb, prefix='feat.'):
    """network.py:53-68.  x [1,3,N], emb [1,32,N] -> [1,1408,N]."""
    n = x.shape[2]
    x1 = _c1(sd, prefix + 'conv1', x)
    e1 = _c1(sd, prefix + 'e_conv1', emb)
    pf1 = torch.cat((x1, e1), dim=1)
    x2 = _c1(sd, prefix + 'conv2', x1)
    e2 = _c1(sd, prefix + 'e_conv2', e1)
    pf2 = torch.cat((x2, e2), dim=1)
    y = _c1(sd, prefix + 'conv5', pf2)
    y = _c1(sd, prefix + 'conv6', y)
    ap = F.avg_pool1d(y, n).view(-1, 1024, 1).repeat(1, 1, n)
    return torch.cat([pf1, pf2, ap], 1)


def posenet_geometry(sd, out_img, x, choose, obj, num_obj):
    """network.py:98-132 with the encoder output given.  out_img [1,32,H,W],
    x [1,N,3], choose [1,1,N] int64, obj [1,1] int64.
    Returns (r [1,N,4], t [1,N,3], c [1,N,1], emb [1,32,N])."""
    bs, di = out_img.shape[:2]
    n = x.shape[1]
    emb = torch.gather(out_img.reshape(bs, di, -1), 2, choose.repeat(1, di, 1)).contiguous()
    ap = posenet_feat(sd, x.transpose(2, 1).contiguous(), emb)
    outs = []
    for h, width in (('r', 4), ('t', 3), ('c', 1)):
        y = _c1(sd, 'conv1_' + h, ap)
        y = _c1(sd, 'conv2_' + h, y)
        y = _c1(sd, 'conv3_' + h, y)
        y = _c1(sd, 'conv4_' + h, y, relu=False)
        if h == 'c':
            y = torch.sigmoid(y)
        y = y.view(bs, num_obj, width, n)
        outs.append(torch.index_select(y[0], 0, obj[0]).transpose(2, 1).contiguous())
    return outs[0], outs[1], outs[2], emb.detach()


def refiner_feat(sd, x, emb, prefix='feat.'):
    """network.py:151-168 -> [1,1024]."""
    n = x.shape[2]
    x1 = _c1(sd, prefix + 'conv1', x)
    e1 = _c1(sd, prefix + 'e_conv1', emb)
    x2 = _c1(sd, prefix + 'conv2', x1)
    e2 = _c1(sd, prefix + 'e_conv2', e1)
    pf3 = torch.cat([x1, e1, x2, e2], dim=1)
    y = _c1(sd, prefix + 'conv5', pf3)
    y = _c1(sd, prefix + 'conv6', y)
    return F.avg_pool1d(y, n).view(-1, 1024)


def refiner_forward(sd, new_points, emb, obj, num_obj):
    """network.py:187-206.  new_points [1,N,3], emb [1,32,N] -> (r [1,4], t [1,3])."""
    ap = refiner_feat(sd, new_points.transpose(2, 1).contiguous(), emb)
    out = []
    for h, width in (('r', 4), ('t', 3)):
        y = F.relu(F.linear(ap, sd['conv1_' + h + '.weight'], sd['conv1_' + h + '.bias']))
        y = F.relu(F.linear(y, sd['conv2_' + h + '.weight'], sd['conv2_' + h + '.bias']))
        y = F.linear(y, sd['conv3_' + h + '.weight'], sd['conv3_' + h + '.bias']).view(1, num_obj, width)
        out.append(torch.index_select(y[0], 0, obj[0]))
    return out[0], out[1]


# ----------------------------------------------------------------------------- pose loops
def _np(t):
    return t.detach().cpu().numpy()


def live_prediction(sd_est, sd_ref, out_img, points, choose, obj, num_obj, refine_calls=2):
    """pipeline/utils.py:564-571: estimator -> get_new_points -> estimator pose ->
    `refine_calls` refiner calls on the SAME new_points (quirk) -> one composition.
    Returns dict(q [4] fp64 wxyz, t [3] fp64, which_max, my_r, my_t, r2, t2)."""
    r, t, c, emb = posenet_geometry(sd_est, out_img, points, choose, obj, num_obj)
    rn, tn, cn, pn = _np(r)[0], _np(t)[0], _np(c)[0, :, 0], _np(points)[0]
    i, my_r, my_t = pose_math.estimator_prediction(rn, tn, cn, pn)
    new_pts = torch.from_numpy(pose_math.new_points(rn, tn, cn, pn))[None]
    r2 = t2 = None
    for _ in range(refine_calls):
        r2, t2 = refiner_forward(sd_ref, new_pts, emb, obj, num_obj)
    if r2 is None:
        return dict(q=my_r.astype(np.float64), t=my_t.astype(np.float64), which_max=i,
                    my_r=my_r, my_t=my_t)
    q, tt = pose_math.refined_prediction(_np(r2), _np(t2), my_r, my_t)
    return dict(q=q, t=tt, which_max=i, my_r=my_r, my_t=my_t, r2=_np(r2)[0], t2=_np(t2)[0])


def canonical_prediction(sd_est, sd_ref, out_img, points, choose, obj, num_obj, iterations=2):
    """DenseFusion/tools/eval_linemod.py:81-114: per iteration the cloud is re-expressed
    in the *composed* pose (fp32 cloud, pose cast to fp32), refiner, fp64 compose."""
    r, t, c, emb = posenet_geometry(sd_est, out_img, points, choose, obj, num_obj)
    rn, tn, cn, pn = _np(r)[0], _np(t)[0], _np(c)[0, :, 0], _np(points)[0]
    i, my_r, my_t = pose_math.estimator_prediction(rn, tn, cn, pn)
    my_r = my_r.astype(np.float64); my_t = my_t.astype(np.float64)
    for _ in range(iterations):
        T = torch.from_numpy(my_t.astype(np.float32)).view(1, 1, 3)            # :92
        M = pose_math.quaternion_matrix(my_r)                                  # :93
        R = torch.from_numpy(M[:3, :3].astype(np.float32)).view(1, 3, 3)       # :94
        M[0:3, 3] = my_t                                                       # :95
        new_pts = torch.bmm(points - T, R).contiguous()                        # :97
        r2, t2 = refiner_forward(sd_ref, new_pts, emb, obj, num_obj)           # :98
        r2n = _np(r2).reshape(4); t2n = _np(t2).reshape(3)
        n2 = np.sqrt((r2n * r2n).sum(dtype=np.float32)).astype(np.float32)     # :100
        r2n = (r2n / n2).astype(np.float32)
        M2 = pose_math.quaternion_matrix(r2n)                                  # :103
        M2[0:3, 3] = t2n                                                       # :104
        Mf = np.dot(M, M2)                                                     # :106
        Rf = Mf.copy(); Rf[0:3, 3] = 0
        my_r = pose_math.quaternion_from_matrix_precise(Rf)                    # :109
        my_t = np.array([Mf[0, 3], Mf[1, 3], Mf[2, 3]])                        # :110
    return dict(q=my_r, t=my_t, which_max=i)


# ----------------------------------------------------------------------------- losses
def _base_from_quat(pred_r):
    """loss.py:17-27 / loss_refiner.py:19-29 : [P,4] normalised -> [P,3,3] (row-major)."""
    q0, q1, q2, q3 = pred_r[:, 0], pred_r[:, 1], pred_r[:, 2], pred_r[:, 3]
    rows = (1.0 - 2.0 * (q2 ** 2 + q3 ** 2), 2.0 * q1 * q2 - 2.0 * q0 * q3, 2.0 * q0 * q2 + 2.0 * q1 * q3,
            2.0 * q1 * q2 + 2.0 * q3 * q0, 1.0 - 2.0 * (q1 ** 2 + q3 ** 2), -2.0 * q0 * q1 + 2.0 * q2 * q3,
            -2.0 * q0 * q2 + 2.0 * q1 * q3, 2.0 * q0 * q1 + 2.0 * q2 * q3, 1.0 - 2.0 * (q1 ** 2 + q2 ** 2))
    return torch.stack(rows, dim=1).view(-1, 3, 3)


def knn_top1_np(ref, query):
    """0-based nearest ref index per query; fp32 ((dx^2+dy^2)+dz^2), lowest index on
    ties -- the arithmetic of knn_cpu.cpp:8-16 with k=1.  ref [N,3], query [M,3]."""
    ref = np.asarray(ref, np.float32); query = np.asarray(query, np.float32)
    out = np.empty(len(query), np.int64)
    for s in range(0, len(query), 1024):
        q = query[s:s + 1024]
        d = None
        for a in range(3):
            diff = ref[None, :, a] - q[:, None, a]
            sq = diff * diff
            d = sq if d is None else d + sq
        out[s:s + 1024] = d.argmin(axis=1)
    return out


def loss_refine(pred_r, pred_t, target, model_points, idx, points, sym_list):
    """loss_refiner.py:12-64 for bs=1.  pred_r [1,4], pred_t [1,3], target/model [1,M,3],
    points [1,N,3].  Returns (dis scalar tensor, new_points [1,N,3], new_target [1,M,3], pred [1,M,3])."""
    m = model_points.shape[1]
    q = pred_r.view(1, 4)
    q = q / torch.norm(q, dim=1, keepdim=True)
    ori_base = _base_from_quat(q)                       # [1,3,3]
    base = ori_base.transpose(2, 1).contiguous()
    t = pred_t.view(1, 1, 3)
    pred = torch.bmm(model_points.view(1, m, 3), base) + t
    tgt = target.view(1, m, 3)
    if int(idx.view(-1)[0]) in sym_list:
        j = knn_top1_np(tgt[0].detach().numpy(), pred[0].detach().numpy())
        tgt_sel = tgt[0][torch.from_numpy(j)].view(1, m, 3)
    else:
        tgt_sel = tgt
    dis = torch.mean(torch.norm(pred - tgt_sel, dim=2), dim=1)
    new_pts = torch.bmm(points.view(1, -1, 3) - t, ori_base).contiguous()
    new_tgt = torch.bmm(tgt - t, ori_base).contiguous()
    return dis, new_pts.detach(), new_tgt.detach(), pred


def loss_estimator(pred_r, pred_t, pred_c, target, model_points, idx, points, w, refine, sym_list):
    """loss.py:12-73 for bs=1.  pred_r [1,N,4], pred_t [1,N,3], pred_c [1,N,1].
    Returns (loss, dis_at_max, new_points, new_target, pred [N,M,3])."""
    n = pred_c.shape[1]
    m = model_points.shape[1]
    q = pred_r.view(n, 4)
    q = q / torch.norm(q, dim=1, keepdim=True)
    ori_base = _base_from_quat(q)
    base = ori_base.transpose(2, 1).contiguous()
    mp = model_points.view(1, m, 3).expand(n, m, 3)
    tgt = target.view(1, m, 3).expand(n, m, 3)
    pt = pred_t.view(n, 1, 3)
    pts = points.view(n, 1, 3)
    c = pred_c.view(n)
    pred = torch.bmm(mp, base) + (pts + pt)
    tgt_used = tgt
    if (not refine) and int(idx.view(-1)[0]) in sym_list:
        j = knn_top1_np(target.view(m, 3).detach().numpy(), pred.detach().reshape(-1, 3).numpy())
        tgt_used = target.view(m, 3)[torch.from_numpy(j)].view(n, m, 3)
    d = torch.norm(pred - tgt_used, dim=2)
    dis = d.mean(dim=1)
    std = d.std(dim=1)
    loss = torch.mean((dis + 2 * std) * c - w * torch.log(c), dim=0)
    i = int(torch.argmax(c))
    tt = (pt[i] + pts[i]).view(1, 1, 3)
    b = ori_base[i].view(1, 3, 3)
    new_pts = torch.bmm(points.view(1, n, 3) - tt, b).contiguous()
    new_tgt = torch.bmm(target.view(1, m, 3) - tt, b).contiguous()
    return loss, dis[i], new_pts.detach(), new_tgt.detach(), pred


# ----------------------------------------------------------------------------- bf16 training emulation
def _bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def refiner_train_bf16_emulation(sd, new_points, emb, obj, num_obj, d_r, d_t):
    """Test infrastructure: PoseRefineNet forward + backward (network.py:151-206 and its autograd) for ONE object in
    fp32 arithmetic with bf16 ROUNDING at exactly the points where the B200 training path stores bf16 (csrc/train.cuh):
    trunk weights conv2/e_conv2/conv5/conv6 and head weights conv1_{r,t}/conv2_{r,t}, the activations x1|e1|x2|e2, h5 and
    the head activations ap/g1/g2, the pooled-layer gradient, and the activation gradients dZ5, dZ2|dZe2, the conv5 share
    of dX1|dE1, dZ1|dZe1 and the head gradients.  conv3_{r,t}, the loss and the first-layer weight gradients stay fp32.  Separates "the
    kernels compute what they are meant to" (tight tolerance against this) from "bf16 differs from fp32" (loose
    tolerance against plain autograd).  new_points [1,N,3], emb [1,32,N], d_r [4], d_t [3] ->
    (r [4], t [3], {state_dict key: gradient})."""
    W = {k: v.detach().to(torch.float32) for k, v in sd.items()}
    w2d = lambda k: W[k].reshape(W[k].shape[0], -1)
    n = new_points.shape[1]
    p = new_points[0].to(torch.float32)                      # [N,3]
    e = emb[0].t().contiguous().to(torch.float32)            # [N,32]
    o = int(obj.reshape(-1)[0])
    x1 = _bf(F.relu(p @ w2d('feat.conv1.weight').t() + W['feat.conv1.bias']))
    e1 = _bf(F.relu(e @ w2d('feat.e_conv1.weight').t() + W['feat.e_conv1.bias']))
    W2, We2, W5, W6 = (_bf(w2d('feat.%s.weight' % k)) for k in ('conv2', 'e_conv2', 'conv5', 'conv6'))
    x2 = _bf(F.relu(x1 @ W2.t() + W['feat.conv2.bias']))
    e2 = _bf(F.relu(e1 @ We2.t() + W['feat.e_conv2.bias']))
    pf = torch.cat([x1, e1, x2, e2], 1)                      # [N,384]
    h5 = _bf(F.relu(pf @ W5.t() + W['feat.conv5.bias']))
    y6 = F.relu(h5 @ W6.t() + W['feat.conv6.bias'])          # never stored: only its sign bits and column mean
    ap = y6.mean(0, keepdim=True)                            # [1,1024]
    g = {}
    r_t, dz1h = [], []
    ap_b = _bf(ap)                                           # heads run on the tensor cores too: bf16 operands, fp32 accumulate
    for h, width, d in (('r', 4, d_r), ('t', 3, d_t)):
        Wa, Wb = _bf(W['conv1_%s.weight' % h]), _bf(W['conv2_%s.weight' % h])
        g1 = _bf(F.relu(ap_b @ Wa.t() + W['conv1_%s.bias' % h]))
        g2 = _bf(F.relu(g1 @ Wb.t() + W['conv2_%s.bias' % h]))
        w3 = W['conv3_%s.weight' % h][o * width:(o + 1) * width]
        r_t.append((g2 @ w3.t() + W['conv3_%s.bias' % h][o * width:(o + 1) * width])[0])
        d = d.reshape(1, width).to(torch.float32)
        gw3 = torch.zeros_like(W['conv3_%s.weight' % h]); gb3 = torch.zeros_like(W['conv3_%s.bias' % h])
        gw3[o * width:(o + 1) * width] = d.t() @ g2; gb3[o * width:(o + 1) * width] = d[0]
        dz2_f = (d @ w3) * (g2 > 0)
        dz2 = _bf(dz2_f)
        dz1_f = (dz2 @ Wb) * (g1 > 0)
        dz1 = _bf(dz1_f)
        g['conv3_%s.weight' % h], g['conv3_%s.bias' % h] = gw3, gb3
        g['conv2_%s.weight' % h], g['conv2_%s.bias' % h] = dz2.t() @ g1, dz2_f[0]
        g['conv1_%s.weight' % h], g['conv1_%s.bias' % h] = dz1.t() @ ap_b, dz1_f[0]
        dz1h.append(dz1 @ Wa)
    g6 = _bf(_bf(dz1h[0] + dz1h[1]) * (1.0 / float(n)))      # dAP stored as bf16, AvgPool1d's 1/N applied where dY6 is formed
    dy6 = (y6 > 0).to(torch.float32) * g6                    # [N,1024]
    g['feat.conv6.weight'], g['feat.conv6.bias'] = dy6.t() @ h5, dy6.sum(0)
    dz5_f = (dy6 @ W6) * (h5 > 0)
    g['feat.conv5.bias'] = dz5_f.sum(0)
    dz5 = _bf(dz5_f)
    g['feat.conv5.weight'] = dz5.t() @ pf
    dpf = dz5 @ W5                                           # [N,384]
    dz22_f = dpf[:, 128:] * (pf[:, 128:] > 0)
    g['feat.conv2.bias'], g['feat.e_conv2.bias'] = dz22_f[:, :128].sum(0), dz22_f[:, 128:].sum(0)
    dz22 = _bf(dz22_f)
    part = _bf(dpf[:, :128])
    g['feat.conv2.weight'] = dz22[:, :128].t() @ x1
    g['feat.e_conv2.weight'] = dz22[:, 128:].t() @ e1
    dz1_f = (dz22[:, :128] @ W2 + part[:, :64]) * (x1 > 0)
    dze1_f = (dz22[:, 128:] @ We2 + part[:, 64:]) * (e1 > 0)
    g['feat.conv1.bias'], g['feat.e_conv1.bias'] = dz1_f.sum(0), dze1_f.sum(0)
    g['feat.conv1.weight'] = _bf(dz1_f).t() @ p
    g['feat.e_conv1.weight'] = _bf(dze1_f).t() @ e
    return r_t[0], r_t[1], {k: v.reshape(sd[k].shape) for k, v in g.items()}
