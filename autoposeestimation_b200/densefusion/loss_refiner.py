"""Drop-in `Loss_refine` (reference: DenseFusion/lib/loss_refiner.py:12-76): ADD / ADD-S distance of one
refinement step plus the cloud / target re-expressed in the predicted frame.

Returns (dis, new_points, new_target, pred) like the reference.  Forward AND backward run in one fused sm_100a kernel
(csrc/knn.cu: refine_loss_kernel -- transform, nearest neighbour for symmetric objects, mean distance, its analytic
gradient w.r.t. the predicted quaternion / translation, and the cloud / target of the next iteration), wrapped in an
autograd node so that `dis.backward()` works for refiner training (train.py:221-222); nearest-neighbour indices are
constants of the backward pass, as in the reference (:45 detaches them).  `add_metric` is the forward-only evaluation
path for batches with shared model clouds.  CUDA tensors only: there is no CPU path."""
import torch
from torch.nn.modules.loss import _Loss

from .. import ops


def quat_to_base(q):
    """[P,4] normalised wxyz -> [P,3,3] rotation, rows as written at loss_refiner.py:19-29."""
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    rows = [1.0 - 2.0 * (y * y + z * z), 2.0 * x * y - 2.0 * w * z, 2.0 * w * y + 2.0 * x * z,
            2.0 * x * y + 2.0 * z * w, 1.0 - 2.0 * (x * x + z * z), -2.0 * w * x + 2.0 * y * z,
            -2.0 * w * y + 2.0 * x * z, 2.0 * w * x + 2.0 * y * z, 1.0 - 2.0 * (x * x + y * y)]
    return torch.stack(rows, dim=1).view(-1, 3, 3)


class _RefineLossFn(torch.autograd.Function):
    """dis (and the next-iteration cloud / target) from the fused kernel; its analytic gradient w.r.t. pred_r / pred_t
    is produced by the same launch (csrc/knn.cu: refine_loss_kernel) and scaled by the incoming gradient here."""

    @staticmethod
    def forward(ctx, pred_r, pred_t, target, model_points, points, sym):
        out = ops.refine_loss(pred_r, pred_t, model_points, target, points, sym)
        ctx.save_for_backward(out['d_r'], out['d_t'])
        ctx.shapes = (pred_r.shape, pred_t.shape)
        ctx.mark_non_differentiable(out['new_points'], out['new_target'])
        return out['dis'], out['new_points'], out['new_target']

    @staticmethod
    def backward(ctx, g_dis, _gp, _gt):
        d_r, d_t = ctx.saved_tensors
        g = g_dis.reshape(-1, 1)
        return (g * d_r).reshape(ctx.shapes[0]), (g * d_t).reshape(ctx.shapes[1]), None, None, None, None


def loss_calculation(pred_r, pred_t, target, model_points, idx, points, num_point_mesh, sym_list):
    """loss_refiner.py:12-64.  The reference handles one object (bs = 1); a leading batch of B objects is accepted
    (pred_r [B,4], pred_t [B,3], target / model_points [B,M,3], idx [B(,1)], points [B,N,3]) and gives dis [B]."""
    if not pred_r.is_cuda:
        raise ops._lib.ApeError('Loss_refine: tensors must be on a CUDA device (no CPU fallback)')
    B = pred_r.reshape(-1, 4).shape[0]
    m = num_point_mesh
    model_points = model_points.reshape(B, m, 3)
    target = target.reshape(B, m, 3)
    points = points.reshape(B, -1, 3)
    sym = None
    if len(sym_list) > 0:
        table = torch.zeros((int(max(sym_list)) + 2,), dtype=torch.uint8, device=pred_r.device)
        table[list(int(v) for v in sym_list)] = 1
        sym = table[idx.reshape(-1).clamp(max=table.numel() - 1)]
    dis, new_points, new_target = _RefineLossFn.apply(pred_r.reshape(B, 4), pred_t.reshape(B, 3), target.detach(),
                                                      model_points.detach(), points.detach(), sym)
    with torch.no_grad():                                            # `pred` (:39) is returned for callers that plot it
        q = pred_r.reshape(B, 4)
        R = quat_to_base(q / torch.norm(q, dim=1, keepdim=True))
        pred = torch.bmm(model_points, R.transpose(2, 1)) + pred_t.reshape(B, 1, 3)
    return dis, new_points, new_target, pred


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
