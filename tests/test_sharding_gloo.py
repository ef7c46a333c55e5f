"""CPU, world_size 2 over gloo: the sharding plumbing of the multi-GPU paths: contiguous ranges,
no data-path collective, scalar summaries all-reduced; and bench.py's max-over-ranks timing reduction."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from autoposeestimation_b200 import sharding


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n_inst, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import clib
    from autoposeestimation_b200 import synthetic as synth
    d = synth.adds_instances(2, n_inst, n_models=3, n_model_pts=200, n_pred_pts=50, n_sym=1)
    from oracle import pose_math as pm

    def dis_fn(lo, hi):                        # CPU stand-in for the per-rank GPU kernel (oracle arithmetic)
        out = []
        for i in range(lo, hi):
            m = d['models'][d['cls'][i]]
            R = pm.quaternion_matrix(d['q_gt'][i])[:3, :3].astype(np.float32)
            tgt = (m @ R.T + d['t_gt'][i]).astype(np.float32)
            sym = bool(d['sym'][d['cls'][i]])
            sub = m[d['subsample']] if sym else m
            out.append(clib.add_metric(d['q_pred'][i], d['t_pred'][i], sub, tgt, sym))
        return np.array(out)

    mean, frac, n = sharding.sharded_add_eval(n_inst, dis_fn)
    lo, hi = sharding.shard_bounds(n_inst)
    t_max = sharding.allreduce_scalars([1.0 + rank], op='max')[0]
    q.put((rank, lo, hi, mean, frac, n, t_max))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_eval_matches_single_process():
    n_inst = 37
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_inst, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    (r0, lo0, hi0, m0, f0, n0, t0), (r1, lo1, hi1, m1, f1, n1, t1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 19, 19, 37) and n0 == n1 == n_inst        # balanced contiguous shards
    assert m0 == m1 and f0 == f1 and t0 == t1 == 2.0                              # identical summaries, MAX timing
    # single-process reference value
    from oracle import clib, pose_math as pm
    from autoposeestimation_b200 import synthetic as synth
    d = synth.adds_instances(2, n_inst, n_models=3, n_model_pts=200, n_pred_pts=50, n_sym=1)
    tot = 0.0
    for i in range(n_inst):
        m = d['models'][d['cls'][i]]
        R = pm.quaternion_matrix(d['q_gt'][i])[:3, :3].astype(np.float32)
        tgt = (m @ R.T + d['t_gt'][i]).astype(np.float32)
        sym = bool(d['sym'][d['cls'][i]])
        tot += clib.add_metric(d['q_pred'][i], d['t_pred'][i], m[d['subsample']] if sym else m, tgt, sym)
    assert abs(m0 - tot / n_inst) < 1e-12


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 64, 100000):
        for w in (1, 2, 3, 8):
            b = [sharding.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def _grad_worker(rank, world, port, B, q):
    """Data-parallel refiner step on CPU: this rank's shard of the batch -> per-sample fp32 gradient (oracle autograd,
    the CPU stand-in for the per-rank GPU backward) -> flat vector -> sharding.allreduce_gradient."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    flat = _shard_gradient(B, *sharding.shard_bounds(B))
    sharding.allreduce_gradient(flat)
    q.put((rank, flat.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def _shard_gradient(B, lo, hi, nobj=2, N=40, M=30):
    from oracle import densefusion as odf
    from autoposeestimation_b200 import synthetic as synth
    sd = {k: torch.from_numpy(np.array(v)).requires_grad_(True) for k, v in synth.refiner_state_dict(3, nobj).items()}
    rng = np.random.RandomState(0)
    pts = torch.from_numpy((rng.randn(B, N, 3) * 0.05).astype(np.float32)); emb = torch.from_numpy(rng.randn(B, 32, N).astype(np.float32))
    idx = torch.from_numpy(rng.randint(0, nobj, (B, 1))); model = torch.from_numpy(((rng.rand(B, M, 3) - 0.5) * 0.2).astype(np.float32))
    for b in range(lo, hi):
        r, t = odf.refiner_forward(sd, pts[b:b + 1], emb[b:b + 1], idx[b:b + 1], nobj)
        d, _, _, _ = odf.loss_refine(r, t, model[b:b + 1] + 0.01, model[b:b + 1], idx[b], pts[b:b + 1], [])
        d.sum().backward()
    return torch.cat([(v.grad if v.grad is not None else torch.zeros_like(v)).reshape(-1) for _, v in sorted(sd.items())]).detach()


def test_two_rank_gradient_allreduce_equals_whole_batch_gradient():
    B = 5
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, B, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=180) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    whole = _shard_gradient(B, 0, B).numpy()
    assert np.array_equal(res[0], res[1])                                        # every rank holds the same reduced vector
    assert np.allclose(res[0], whole, rtol=1e-5, atol=1e-7) and np.abs(whole).max() > 0   # SUM over shards (train.py:222 semantics)
