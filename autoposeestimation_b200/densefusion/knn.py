"""Drop-in `KNearestNeighbor` (reference: DenseFusion/lib/knn/__init__.py:9-23).

The reference class is a legacy autograd.Function instance that torch >= 1.5 refuses to call; the
drop-in keeps the constructor / call syntax as a plain callable.  Indices are 1-based int64 [B,k,M],
non-differentiable (callers `.detach()` them, loss.py:45).  Like the reference it moves its inputs to the
GPU (`.float().cuda()`, :16-17); there is no CPU implementation."""
import torch

from .. import ops


class KNearestNeighbor:
    def __init__(self, k, arith=ops.KNN_ARITH_CPU):
        self.k = k
        self.arith = arith          # APE_KNN_ARITH_CPU: knn_cpu.cpp arithmetic; APE_KNN_ARITH_FMA: knn.cu arithmetic

    def forward(self, ref, query):
        ref = ref.detach().float().cuda()
        query = query.detach().float().cuda()
        return ops.knn(ref, query, self.k, self.arith)

    __call__ = forward
