"""GPU: the embedding hand-over (network.py:100-102 `torch.gather(emb, 2, choose)`) -- stand-alone gather kernel on device
and zero-copy pinned maps, channels-last maps, the pre-gathered entry points and the host-buffer Runner: every path must
give bit-identical poses (it moves the same fp32 values)."""
import numpy as np
import pytest
import torch

from oracle import synth

pytestmark = pytest.mark.gpu


def _inputs(seed, B, N, hw=(40, 60), nobj=3):
    out_img, cloud, choose, idx = synth.posenet_inputs(seed, N, hw, nobj, batch=B)
    return [torch.from_numpy(np.ascontiguousarray(a)) for a in (out_img, cloud, choose, idx)]


def _ref_gather(out_img, choose):
    B, N = choose.shape[0], choose.shape[-1]
    return torch.gather(out_img.reshape(B, 32, -1), 2, choose.reshape(B, 1, N).expand(B, 32, N))


@pytest.mark.parametrize('N', [1, 100, 500, 1000])
def test_gather_emb_device_and_zero_copy(N):
    from autoposeestimation_b200 import ops
    out_img, cloud, choose, idx = _inputs(3, 5, N)
    want = _ref_gather(out_img, choose)
    ch = choose.cuda()
    assert torch.equal(ops.gather_emb(out_img.cuda(), ch).cpu(), want)                       # device map, NCHW
    assert torch.equal(ops.gather_emb(out_img.pin_memory(), ch).cpu(), want)                 # pinned host map read in place
    cl = out_img.contiguous(memory_format=torch.channels_last)
    assert torch.equal(ops.gather_emb(cl.cuda().contiguous(memory_format=torch.channels_last), ch).cpu(), want)
    assert torch.equal(ops.gather_emb(cl.pin_memory(), ch).cpu(), want)                      # zero-copy, one 128-byte line per point
    with pytest.raises(ops._lib.ApeError):
        ops.gather_emb(out_img, ch)                                                          # pageable host memory: loud error


def test_gather_clamps_out_of_range_indices():
    from autoposeestimation_b200 import ops
    out_img, cloud, choose, idx = _inputs(4, 2, 64)
    bad = choose.clone(); bad[0, 0, 0] = -5; bad[1, 0, 3] = 10 ** 9
    got = ops.gather_emb(out_img.cuda(), bad.cuda()).cpu()
    want = _ref_gather(out_img, bad.clamp(0, out_img.shape[2] * out_img.shape[3] - 1))
    assert torch.equal(got, want)


def test_pipeline_layouts_bitwise_equal():
    """NCHW map, channels-last map and the pre-gathered embedding give the same bits (PoseNet + 2 refine iterations)."""
    from autoposeestimation_b200 import ops
    nobj, B, N = 3, 4, 500
    est = ops.NetHandle(ops.NET_POSENET, synth.posenet_state_dict(5, nobj), nobj, B, N)
    ref = ops.NetHandle(ops.NET_REFINER, synth.refiner_state_dict(1005, nobj), nobj, B, N)
    out_img, cloud, choose, idx = [t.cuda() for t in _inputs(5, B, N, nobj=nobj)]
    p0, w0 = ops.pose_pipeline(est, ref, out_img, cloud, choose, idx)
    cl = out_img.contiguous(memory_format=torch.channels_last)
    p1, w1 = ops.pose_pipeline(est, ref, cl, cloud, choose, idx)
    emb = ops.gather_emb(out_img, choose)
    p2, w2 = ops.pose_pipeline(est, ref, emb, cloud, None, idx, gathered=True)
    assert torch.equal(p0, p1) and torch.equal(p0, p2) and torch.equal(w0, w1) and torch.equal(w0, w2)
    r0 = est.posenet_forward(out_img, cloud, choose, idx)
    r1 = est.posenet_forward(cl, cloud, choose, idx)
    r2 = est.posenet_forward(emb, cloud, None, idx, gathered=True)
    for a, b, c in zip(r0, r1, r2):
        assert torch.equal(a, b) and torch.equal(a, c)


@pytest.mark.parametrize('graph', [False, True])
def test_runner_transfer_modes_agree(graph):
    """Host-buffer API: whole-map copy (round 1), split zero-copy + host-pool gather, host pool alone (pageable map),
    channels-last zero-copy, device-resident map -- identical poses, over enough submits to reuse both slots (and graphs)."""
    from autoposeestimation_b200 import ops
    from autoposeestimation_b200.densefusion.estimate_poses import Runner
    nobj, B, N, hw = 3, 6, 300, (40, 60)
    est = ops.NetHandle(ops.NET_POSENET, synth.posenet_state_dict(6, nobj), nobj, B, N)
    ref = ops.NetHandle(ops.NET_REFINER, synth.refiner_state_dict(1006, nobj), nobj, B, N)
    batches = [_inputs(20 + i, B, N, hw, nobj) for i in range(3)]
    want = []
    for b in batches:
        d = [t.cuda() for t in b]
        want.append(ops.pose_pipeline(est, ref, *d)[0].cpu())

    def run(prep, **kw):
        r = Runner(est, ref, B, N, hw[0] * hw[1], use_graph=graph, **kw)
        outs = []
        for rep in range(2):
            for b in batches:
                r.submit(*prep(b))
                outs.append(r.drain())
        return outs, r

    pin = lambda b: [t.pin_memory() for t in b]
    for name, prep, kw in (('full', pin, dict(transfer='full')),
                           ('split', pin, dict(zero_copy_fraction=0.5)),
                           ('three-way', pin, dict(zero_copy_fraction=0.34, dma_fraction=0.33)),
                           ('copy engine only', pin, dict(zero_copy_fraction=0.0, dma_fraction=1.0)),
                           ('auto', pin, {}),
                           ('zero-copy only', pin, dict(zero_copy_fraction=1.0)),
                           ('pageable', lambda b: list(b), {}),
                           ('channels_last', lambda b: [b[0].contiguous(memory_format=torch.channels_last).pin_memory()] + [t.pin_memory() for t in b[1:]], {}),
                           ('device', lambda b: [b[0].cuda()] + [t.pin_memory() for t in b[1:]], {})):
        outs, r = run(prep, **kw)
        for i, o in enumerate(outs):
            assert torch.equal(o, want[i % 3]), (name, i)
        if name == 'auto':
            c = r.calibration
            assert c is not None and abs(c['zero_copy_fraction'] + c['dma_fraction'] + c['host_fraction'] - 1.0) < 1e-6
