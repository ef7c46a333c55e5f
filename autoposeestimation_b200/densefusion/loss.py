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
    knn = KNearestNeighbor(1)
    bs, num_p, _ = pred_c.size()
    m = num_point_mesh
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


class Loss(_Loss):
    def __init__(self, num_points_mesh, sym_list):
        super().__init__(True)
        self.num_pt_mesh = num_points_mesh
        self.sym_list = sym_list

    def forward(self, pred_r, pred_t, pred_c, target, model_points, idx, points, w, refine):
        return loss_calculation(pred_r, pred_t, pred_c, target, model_points, idx, points, w, refine,
                                self.num_pt_mesh, self.sym_list)
