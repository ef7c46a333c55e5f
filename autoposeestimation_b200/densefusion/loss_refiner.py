"""Drop-in `Loss_refine` (reference: DenseFusion/lib/loss_refiner.py:12-76): ADD / ADD-S distance of one
refinement step plus the cloud / target re-expressed in the predicted frame.

Returns (dis, new_points, new_target, pred) like the reference.  The nearest-neighbour search for
symmetric objects is the sm_100a kNN kernel (indices are not differentiable, as in the reference); the
remaining arithmetic is a handful of differentiable torch ops on the device so that `dis.backward()`
works for refiner training (train.py:221-222).  `add_metric` gives the fused, forward-only evaluation
path (transform + kNN + gather + mean in one kernel) used by evaluation drivers."""
import torch
from torch.nn.modules.loss import _Loss

from .. import ops
from .knn import KNearestNeighbor


def quat_to_base(q):
    """[P,4] normalised wxyz -> [P,3,3] rotation, rows as written at loss_refiner.py:19-29."""
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    rows = [1.0 - 2.0 * (y * y + z * z), 2.0 * x * y - 2.0 * w * z, 2.0 * w * y + 2.0 * x * z,
            2.0 * x * y + 2.0 * z * w, 1.0 - 2.0 * (x * x + z * z), -2.0 * w * x + 2.0 * y * z,
            -2.0 * w * y + 2.0 * x * z, 2.0 * w * x + 2.0 * y * z, 1.0 - 2.0 * (x * x + y * y)]
    return torch.stack(rows, dim=1).view(-1, 3, 3)


def loss_calculation(pred_r, pred_t, target, model_points, idx, points, num_point_mesh, sym_list):
    knn = KNearestNeighbor(1)
    q = pred_r.reshape(1, 4)
    q = q / torch.norm(q, dim=1, keepdim=True)
    R = quat_to_base(q)                                              # ori_base
    t = pred_t.reshape(1, 1, 3)
    model_points = model_points.reshape(1, num_point_mesh, 3)
    target = target.reshape(1, num_point_mesh, 3)
    pred = torch.bmm(model_points, R.transpose(2, 1)) + t            # :39
    tgt = target
    if int(idx.reshape(-1)[0]) in sym_list:                          # :41-47
        inds = knn(target[0].t().unsqueeze(0), pred[0].t().unsqueeze(0)).view(-1) - 1
        tgt = target[:, inds.to(target.device), :]
    dis = torch.mean(torch.norm(pred - tgt, dim=2), dim=1)           # :49
    new_points = torch.bmm(points.reshape(1, -1, 3) - t, R).contiguous()
    new_target = torch.bmm(target - t, R).contiguous()
    return dis, new_points.detach(), new_target.detach(), pred


class Loss_refine(_Loss):
    def __init__(self, num_points_mesh, sym_list):
        super().__init__(True)
        self.num_pt_mesh = num_points_mesh
        self.sym_list = sym_list

    def forward(self, pred_r, pred_t, target, model_points, idx, points):
        return loss_calculation(pred_r, pred_t, target, model_points, idx, points, self.num_pt_mesh, self.sym_list)


def add_metric(pred_r, pred_t, model_points, target, symmetric):
    """Fused forward-only ADD / ADD-S for batches (csrc/knn.cu:add_metric_kernel): pred_r [B,4], pred_t [B,3],
    model_points [B,M,3] or [M,3], target [B,Nt,3] or [Nt,3], symmetric [B] -> dis [B]."""
    return ops.add_metric(pred_r, pred_t, model_points, target, symmetric)
