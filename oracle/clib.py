"""ctypes access to the oracle's C restatement and to the reference's own kNN
binaries under oracle/_ref (TEST INFRASTRUCTURE; see oracle/__init__.py)."""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_i64p = ctypes.POINTER(ctypes.c_int64)
_f32p = ctypes.POINTER(ctypes.c_float)


def build(verbose=False):
    """Compile oracle/_build/liboracle.so and, if /root/reference exists, oracle/_ref/*."""
    r = subprocess.run(['make', '-C', _HERE], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError('oracle build failed')


def _load(rel):
    path = os.path.join(_HERE, rel)
    if not os.path.exists(path):
        return None
    return ctypes.CDLL(path)


_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        _oracle = _load('_build/liboracle.so')
        if _oracle is None:
            build()
            _oracle = _load('_build/liboracle.so')
        _oracle.oracle_knn.argtypes = [_f32p, _f32p, _i64p] + [ctypes.c_int] * 6
        _oracle.oracle_knn.restype = ctypes.c_int
        _oracle.oracle_add_metric.argtypes = [_f32p, _f32p, _f32p, ctypes.c_int, _f32p, ctypes.c_int, ctypes.c_int]
        _oracle.oracle_add_metric.restype = ctypes.c_double
    return _oracle


def _fp(a):
    return a.ctypes.data_as(_f32p)


def knn(ref, query, k=1, mode=0):
    """ref [B,D,N], query [B,D,M] fp32 -> idx [B,k,M] int64, 1-based (oracle restatement)."""
    ref = np.ascontiguousarray(ref, np.float32); query = np.ascontiguousarray(query, np.float32)
    B, D, N = ref.shape; M = query.shape[2]
    idx = np.empty((B, k, M), np.int64)
    rc = oracle_lib().oracle_knn(_fp(ref), _fp(query), idx.ctypes.data_as(_i64p), B, D, N, M, k, mode)
    assert rc == 0
    return idx


def add_metric(quat, t, model, target, symmetric):
    quat = np.ascontiguousarray(quat, np.float32); t = np.ascontiguousarray(t, np.float32)
    model = np.ascontiguousarray(model, np.float32); target = np.ascontiguousarray(target, np.float32)
    return oracle_lib().oracle_add_metric(_fp(quat), _fp(t), _fp(model), len(model), _fp(target), len(target),
                                          int(bool(symmetric)))


def ref_knn_cpu(ref, query, k=1):
    """The reference's own knn_cpu.cpp (compiled unmodified).  None if oracle/_ref is absent."""
    lib = _load('_ref/libknn_cpu_ref.so')
    if lib is None:
        return None
    ref = np.ascontiguousarray(ref, np.float32); query = np.ascontiguousarray(query, np.float32)
    B, D, N = ref.shape; M = query.shape[2]
    idx = np.empty((B, k, M), np.int64)
    lib.ref_knn_cpu.argtypes = [_f32p, _f32p, _i64p] + [ctypes.c_int] * 5
    lib.ref_knn_cpu(_fp(ref), _fp(query), idx.ctypes.data_as(_i64p), B, D, N, M, k)
    return idx


def ref_knn_cuda_lib():
    """The reference's own knn.cu (compiled unmodified for sm_100a) + launcher shim, or None."""
    lib = _load('_ref/libknn_cuda_ref.so')
    if lib is not None:
        lib.ref_knn_cuda.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 5 + [ctypes.c_void_p]
        lib.ref_knn_cuda.restype = ctypes.c_int
    return lib
