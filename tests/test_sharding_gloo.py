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
