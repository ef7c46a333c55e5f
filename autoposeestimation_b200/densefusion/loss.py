"""Drop-in `Loss` (reference: DenseFusion/lib/loss.py:12-85): per-point candidate poses, ADD / ADD-S with the fork's
`(dis + 2 std) c - w log c` objective, and the cloud / target in the frame of the most confident point.  Returns the
reference's 5-tuple (loss, dis, new_points, new_target, pred).

Everything -- the candidate transforms, the nearest-neighbour targets of symmetric objects, the mean / std distances, the
objective, its GRADIENT w.r.t. pred_r / pred_t / pred_c, the arg-max and the re-expressed clouds -- runs in two sm_100a
kernels (csrc/knn.cu: estimator_loss_kernel, estimator_loss_finish_kernel) behind one autograd node; there is no torch
arithmetic on this path.  The reference materialises pred [N,M,3], repeats model / target N times and runs an N*M-query kNN
with an (N*M) x M distance matrix (4 GB at 1000 x 1000, knn.h:33); here `pred` is written once only because it is part of
the returned tuple.  As in the reference the kNN indices are constants of the backward pass (`inds.detach()`, :45),
`new_points` / `new_target` are detached (:73); `dis` and `pred` are returned as values (train.py back-propagates `loss`
only, :222 / :216).  The reference's batch size is 1 (`idx[0]`, `which_max[0]`, :41, :60): other batch sizes raise."""
import torch
from torch.nn.modules.loss import _Loss

from .. import ops


class _EstimatorLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred_r, pred_t, pred_c, target, model_points, points, w, symmetric, want_grad):
        out = ops.estimator_loss(pred_r, pred_t, pred_c, points, model_points, target, symmetric, w, want_grad=want_grad)
        ctx.shapes = (pred_r.shape, pred_t.shape, pred_c.shape)
        if want_grad:
            ctx.save_for_backward(out['d_r'], out['d_t'], out['d_c'])
        dis_best, newp, newt, pred = out['dis_best'].clone(), out['new_points'], out['new_target'], out['pred']
        ctx.mark_non_differentiable(dis_best, newp, newt, pred)
        return out['loss'].clone(), dis_best, newp, newt, pred

    @staticmethod
    def backward(ctx, g_loss, *unused):
        d_r, d_t, d_c = ctx.saved_tensors
        sr, st, sc = ctx.shapes
        return (g_loss * d_r).reshape(sr), (g_loss * d_t).reshape(st), (g_loss * d_c).reshape(sc), None, None, None, None, None, None


def loss_calculation(pred_r, pred_t, pred_c, target, model_points, idx, points, w, refine, num_point_mesh, sym_list):
    if not pred_r.is_cuda:
        raise ops._lib.ApeError('Loss: tensors must be on a CUDA device (no CPU fallback)')
    bs, num_p, _ = pred_c.size()
    if bs != 1:
        raise NotImplementedError('Loss: the reference evaluates one object per call (batch size 1: loss.py:41, :60, train.py:216)')
    symmetric = (not refine) and int(idx.reshape(-1)[0]) in sym_list                        # :40-41
    want_grad = torch.is_grad_enabled() and (pred_r.requires_grad or pred_t.requires_grad or pred_c.requires_grad)
    return _EstimatorLossFn.apply(pred_r, pred_t, pred_c, target, model_points, points, float(w), symmetric, want_grad)


class Loss(_Loss):
    def __init__(self, num_points_mesh, sym_list):
        super().__init__(True)
        self.num_pt_mesh = num_points_mesh
        self.sym_list = sym_list

    def forward(self, pred_r, pred_t, pred_c, target, model_points, idx, points, w, refine):
        return loss_calculation(pred_r, pred_t, pred_c, target, model_points, idx, points, w, refine,
                                self.num_pt_mesh, self.sym_list)
