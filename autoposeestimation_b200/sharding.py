"""Multi-GPU plumbing for the paths that shard: frames / objects / instances are independent, so
every rank takes a contiguous range and there is NO data-path collective; only the final scalar summaries
(counts, sums) are all-reduced.  One process per GPU, torch.distributed (NCCL on the GPU box, gloo in the CPU
tests)."""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_items, rank=None, world_size=None):
    """Contiguous, balanced range [lo, hi) of `n_items` for `rank` (first n % world ranks get one extra)."""
    if rank is None:
        rank, world_size = world()
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_scalars(values, op='sum', device=None):
    """All-reduce a short list of python floats (summary statistics only); identity when not distributed."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device or 'cpu')
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == 'nccl' and t.device.type != 'cuda':
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM if op == 'sum' else dist.ReduceOp.MAX)
    return t.cpu().tolist()


def allreduce_gradient(flat, average=False):
    """The one data-path collective of the scope: sum the refiner's flat fp32 gradient vector over ranks,
    in place (NCCL over NVLink on the GPU box, gloo in the CPU tests).  The reference accumulates the SUM of per-sample
    gradients before optimizer.step() (train.py:222, :231-233), so the data-parallel equivalent is a sum, not a mean;
    `average` divides by the world size for callers that want the mean."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if average:
            flat.div_(dist.get_world_size())
    return flat


def sharded_add_eval(n_instances, dis_fn, threshold=0.02):
    """BASELINE config 3 driver: `dis_fn(lo, hi)` returns the ADD/ADD-S distances (1-D tensor/array) of instances
    [lo, hi) computed on this rank's GPU.  Returns the global (mean distance, fraction below `threshold`,
    n_instances) -- experiments/eval.py:80-84 counts `dis < 0.02`."""
    lo, hi = shard_bounds(n_instances)
    d = dis_fn(lo, hi)
    d = torch.as_tensor(d, dtype=torch.float64).reshape(-1)
    s, c, n = allreduce_scalars([float(d.sum()), float((d < threshold).sum()), float(d.numel())])
    return s / max(n, 1.0), c / max(n, 1.0), int(n)
