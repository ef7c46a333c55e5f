"""Drop-in `Loss` (reference: DenseFusion/lib/loss.py:12-85): per-point candidate poses, ADD / ADD-S with the
fork's `(dis + 2 std) c - w log c` objective, and the cloud / target in the frame of the most confident point.
Returns the reference's 5-tuple (loss, dis, new_points, new_target, pred).  kNN (symmetric objects, not-refine
branch) runs on the sm_100a kernel; the rest is differentiable torch glue on the device."""
import torch
from torch.nn.modules.loss import _Loss

from .. import ops
from .knn import KNearestNeighbor
from .loss_refiner import quat_to_base


def loss_calculation(pred_r, pred_t, pred_c, target, model_points, idx, points, w, refine, num_point_mesh, sym_list):
    if not pred_r.is_cuda:
        raise ops._lib.ApeError('Loss: tensors must be on a CUDA device (no CPU fallback)')
    bs, num_p, _ = pred_c.size()
    m = num_point_mesh
    if bs == 1 and m <= 2048 and not (torch.is_grad_enabled() and (pred_r.requires_grad or pred_t.requires_grad or pred_c.requires_grad)):
        return _fused_forward(pred_r, pred_t, pred_c, target, model_points, idx, points, w, refine, m, sym_list)
    knn = KNearestNeighbor(1)
    q = pred_r.reshape(bs * num_p, 4)
    q = q / torch.norm(q, dim=1, keepdim=True)
    ori_base = quat_to_base(q)
    mp = model_points.reshape(bs, 1, m, 3).expand(bs, num_p, m, 3).reshape(bs * num_p, m, 3)
    tg = target.reshape(bs, 1, m, 3).expand(bs, num_p, m, 3).reshape(bs * num_p, m, 3)
    pt = pred_t.reshape(bs * num_p, 1, 3)
    ps = points.reshape(bs * num_p, 1, 3)
    c = pred_c.reshape(bs * num_p)
    pred = torch.bmm(mp, ori_base.transpose(2, 1)) + (ps + pt)                              # :38
    tg_used = tg
    if (not refine) and int(idx.reshape(-1)[0]) in sym_list:                                 # :40-47
        t0 = target.reshape(bs, m, 3)[0]
        inds = knn(t0.t().unsqueeze(0), pred.reshape(-1, 3).t().unsqueeze(0)).view(-1) - 1
        tg_used = t0[inds.to(t0.device)].view(bs * num_p, m, 3)
    d = torch.norm(pred - tg_used, dim=2)
    dis, std = d.mean(dim=1), d.std(dim=1)
    loss = torch.mean((dis + 2 * std) * c - w * torch.log(c), dim=0)                         # :53
    which = torch.argmax(c.view(bs, num_p), dim=1)[0]
    t = (pt[which] + ps[which]).view(1, 1, 3)
    b = ori_base[which].view(1, 3, 3)
    new_points = torch.bmm(points.reshape(1, bs * num_p, 3) - t, b).contiguous()
    new_target = torch.bmm(tg[0].view(1, m, 3) - t, b).contiguous()
    return loss, dis.view(bs, num_p)[0][which], new_points.detach(), new_target.detach(), pred


def _fused_forward(pred_r, pred_t, pred_c, target, model_points, idx, points, w, refine, m, sym_list):
    """Forward-only path (evaluation, and the refine phase of train.py:205-216 where `loss` is not back-propagated): the
    per-candidate mean / std distances come from ONE fused kernel (csrc/knn.cu: add_metric_kernel with B = num_p candidate
    poses, shared model / target) -- no [N,M,3] gather, no N*M-query kNN, no N x M distance matrix."""
    num_p = pred_c.shape[1]
    with torch.no_grad():
        q = pred_r.reshape(num_p, 4)
        trans = (points.reshape(num_p, 3) + pred_t.reshape(num_p, 3)).contiguous()
        sym = torch.full((num_p,), 1 if ((not refine) and int(idx.reshape(-1)[0]) in sym_list) else 0, dtype=torch.uint8, device=q.device)
        dis, std = ops.add_metric_std(q, trans, model_points.reshape(m, 3), target.reshape(m, 3), sym)
        c = pred_c.reshape(num_p)
        loss = torch.mean((dis + 2 * std) * c - w * torch.log(c), dim=0)                        # :53
        which = torch.argmax(c)
        qn = q / torch.norm(q, dim=1, keepdim=True)
        ori_base = quat_to_base(qn)
        t = trans[which].view(1, 1, 3)
        b = ori_base[which].view(1, 3, 3)
        new_points = torch.bmm(points.reshape(1, num_p, 3) - t, b).contiguous()
        new_target = torch.bmm(target.reshape(1, m, 3) - t, b).contiguous()
        # `pred` (:38) is part of the returned tuple; it is cheap to form once the distances no longer depend on it
        pred = torch.bmm(model_points.reshape(1, m, 3).expand(num_p, m, 3), ori_base.transpose(2, 1)) + trans.view(num_p, 1, 3)
    return loss, dis[which], new_points, new_target, pred


class Loss(_Loss):
    def __init__(self, num_points_mesh, sym_list):
        super().__init__(True)
        self.num_pt_mesh = num_points_mesh
        self.sym_list = sym_list

    def forward(self, pred_r, pred_t, pred_c, target, model_points, idx, points, w, refine):
        return loss_calculation(pred_r, pred_t, pred_c, target, model_points, idx, points, w, refine,
                                self.num_pt_mesh, self.sym_list)
