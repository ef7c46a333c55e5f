"""Drop-in for DenseFusion/tools/utils.py: same names, arguments and return values (numpy on the host,
quaternion [w,x,y,z], metres), computed by the pose-math kernels (csrc/pose_math.cu)."""
import numpy as np
import torch

from .. import ops


def my_estimator_prediction(pred_r, pred_t, pred_c, num_points, bs, cloud):
    """tools/utils.py:7-18 -> (my_pred np[7], my_r np[4], my_t np[3]) of batch element 0."""
    out = ops.pose_select(pred_r.reshape(bs, num_points, 4), pred_t.reshape(bs, num_points, 3), pred_c.reshape(bs, num_points),
                          cloud.reshape(bs, num_points, 3), want_new_points=False)
    my_r = out['my_r'][0].cpu().numpy()
    my_t = out['my_t'][0].cpu().numpy()
    return np.append(my_r, my_t), my_r, my_t


def my_refined_prediction(pred_r, pred_t, my_r, my_t):
    """tools/utils.py:20-40 -> (my_pred np[7] fp64, my_r np[4], my_t np[3])."""
    dev = pred_r.device
    pose_in = torch.from_numpy(np.concatenate([np.asarray(my_r, np.float64), np.asarray(my_t, np.float64)])[None]).to(dev)
    out = ops.pose_compose(pose_in, pred_r.reshape(1, 4), pred_t.reshape(1, 3))[0].cpu().numpy()
    return out.copy(), out[:4].copy(), out[4:].copy()


def get_new_points(pred_r, pred_t, pred_c, points):
    """tools/utils.py:43-86 -> new_points [1,N,3] (detached, on the device).  Batched inputs return [B,N,3]."""
    bs, num_p = pred_c.shape[0], pred_c.shape[1]
    out = ops.pose_select(pred_r, pred_t, pred_c.reshape(bs, num_p), points.reshape(bs, num_p, 3))
    return out['new_points'].detach()
