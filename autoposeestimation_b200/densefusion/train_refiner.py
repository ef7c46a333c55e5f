"""PoseRefineNet refine-phase training step (reference: DenseFusion/tools/train.py:205-238) on the sm_100a kernels.

The reference trains with DataLoader batch 1 and gradient accumulation: per sample and per refinement iteration
`pred_r, pred_t = refiner(new_points, emb, idx)`; `dis, new_points, new_target, _ = criterion_refine(...)`;
`dis.backward()` (train.py:218-222), then one `optimizer.step()` per `opt.batch_size` samples (:231-233) -- i.e. the
applied gradient is the SUM over samples and iterations.  `RefinerTrainer.train_step` computes exactly that sum for a
whole batch of objects at once:

    for ite in range(iterations):                       # train.py:218
        r, t      = trainer forward (bf16 tcgen05 trunk, fp32 heads)            ape_refiner_trainer_forward
        dis, d_r, d_t, new_points, new_target = Loss_refine fwd + bwd           ape_refine_loss
        grads    += backward(d_r, d_t)                                           ape_refiner_trainer_backward
    grads = all_reduce_sum(grads)                       # NCCL over NVLink, the only collective of the scope
    Adam(params, grads)                                 # train.py:149 / :410    ape_adam_step

Multi-GPU is data parallel: every rank holds the full 1.93 M-parameter model, takes its shard of the objects, and
the flat gradient (one 7.7 MB fp32 buffer) is all-reduced once per step.  There is no CPU fallback.
"""
import torch

from .. import ops, sharding


class RefinerTrainer:
    def __init__(self, state_dict, num_obj, max_batch, max_points, sym_list=(), lr=1e-4, betas=(0.9, 0.999), eps=1e-8,
                 iterations=2, device=None):
        """state_dict: PoseRefineNet parameters with the reference's names / shapes (network.py:139-183)."""
        self.h = ops.RefinerTrainerHandle(state_dict, num_obj, max_batch, max_points, device=device)
        self.exp_avg = torch.zeros_like(self.h.params)
        self.exp_avg_sq = torch.zeros_like(self.h.params)
        self.sym_list = set(int(s) for s in sym_list)
        self.lr, self.betas, self.eps, self.iterations = lr, betas, eps, iterations
        self.step_count = 0
        self.num_obj = num_obj
        self._side = None

    # -- pieces (also used by the autograd drop-in in network.py)
    def zero_grad(self):
        self.h.grads.zero_()

    def symmetric_flags(self, idx):
        if not self.sym_list:
            return None
        table = torch.zeros((self.num_obj,), dtype=torch.uint8, device=idx.device)
        table[list(self.sym_list)] = 1
        return table[idx.reshape(-1)]

    def accumulate(self, points, emb, idx, target, model_points):
        """`iterations` x (forward -> Loss_refine -> backward) for a batch; gradients accumulate into the flat vector.
        points [B,N,3] (the estimator's new_points), emb [B,32,N], idx [B], target / model_points [B,M,3].
        Returns dis [iterations, B] (device)."""
        sym = self.symmetric_flags(idx)
        dis_all = []
        for _ in range(self.iterations):
            r, t = self.h.forward(points, emb, idx)
            out = ops.refine_loss(r, t, model_points, target, points, sym)
            self.h.backward(points, emb, idx, out['d_r'], out['d_t'])
            points, target = out['new_points'], out['new_target']          # train.py:223 (detached by construction)
            dis_all.append(out['dis'])
        return torch.stack(dis_all)

    def allreduce_gradient(self):
        """Sum the flat gradient over ranks (NCCL on the GPU box; identity for a single process)."""
        sharding.allreduce_gradient(self.h.grads)

    def allreduce_gradient_overlapped(self):
        """After `self.h.step(...)`: all-reduce the tail block of the gradient (conv6 + heads, 89 %; final once the last
        iteration's conv6 weight gradient is written) on a side stream while the backward pass still walks conv5 / conv2 /
        conv1, the head block on the main stream afterwards; the streams are joined before the optimizer."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.h.device)
        main = torch.cuda.current_stream(self.h.device)
        lo = self.h.wait_bulk(self._side)
        with torch.cuda.stream(self._side):
            dist.all_reduce(self.h.grads[lo:], op=dist.ReduceOp.SUM)
        dist.all_reduce(self.h.grads[:lo], op=dist.ReduceOp.SUM)
        main.wait_stream(self._side)

    def optimizer_step(self):
        self.step_count += 1
        self.h.adam(self.exp_avg, self.exp_avg_sq, self.step_count, self.lr, self.betas, self.eps)

    def train_step(self, points, emb, idx, target, model_points, overlap=True):
        """One optimizer step over this rank's shard of the batch (train.py:215-233).  Returns dis [iterations, B].
        Host calls per step: the whole accumulation phase (ape_refiner_trainer_step), the NCCL all-reduce of the flat
        gradient in two blocks (the tail block overlapped with the end of the backward pass), Adam + weight refresh -- so
        the step stays GPU-bound when the per-rank batch gets small."""
        dis = self.h.step(points, emb, idx, target, model_points, self.symmetric_flags(idx), self.iterations, zero_grad=True)
        if overlap:
            self.allreduce_gradient_overlapped()
        else:
            self.allreduce_gradient()
        self.optimizer_step()
        return dis

    def state_dict(self):
        return self.h.state_dict()

    def load_state_dict(self, sd):
        self.h.load_state_dict(sd)
