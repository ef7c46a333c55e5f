"""ctypes binding of libape_b200.so (the C-ABI in include/ape_b200.h).

The product path has no CPU fallback: if the shared library is missing, or a call is made
without a CUDA device, this module raises -- it never routes through oracle/ or torch eager.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libape_b200.so')

c_int, c_i64, c_dbl, c_vp, c_sz = ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p, ctypes.c_size_t

# name -> (restype, argtypes); every symbol declared in include/ape_b200.h
SIGNATURES = {
    'ape_version': (c_int, []),
    'ape_last_error': (ctypes.c_char_p, []),
    'ape_launch_count': (ctypes.c_uint64, []),
    'ape_profile_enable': (c_int, [c_int]),
    'ape_profile_report': (c_int, [ctypes.c_char_p, c_int]),
    'ape_backproject_choose': (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp]),
    'ape_mask_bbox_choose': (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'ape_surface_work_bytes': (c_sz, [c_int, c_int, c_int]),
    'ape_surface_backproject': (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_int,
                                        c_vp, c_vp, c_vp, c_vp, c_vp]),
    'ape_knn': (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp]),
    'ape_add_metric': (c_int, [c_vp, c_vp, c_vp, c_i64, c_int, c_vp, c_i64, c_int, c_vp, c_int, c_vp, c_vp, c_vp]),
    'ape_add_metric_std': (c_int, [c_vp, c_vp, c_vp, c_i64, c_int, c_vp, c_i64, c_int, c_vp, c_int, c_vp, c_vp, c_vp]),
    'ape_icp_work_bytes': (c_sz, [c_int, c_int]),
    'ape_reconstruct_work_bytes': (c_sz, [c_int]),
    'ape_reconstruct_run': (c_int, [c_vp, c_vp, c_int, c_dbl, c_dbl, c_dbl, c_dbl, c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'ape_icp_p2p': (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_dbl, c_dbl, c_dbl, c_int,
                            c_vp, c_vp, c_vp, c_vp, c_vp]),
    'ape_icp_p2p_ex': (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_dbl, c_dbl, c_dbl, c_int,
                               c_vp, c_vp, c_vp, c_vp, c_vp]),
    'ape_surface_backproject_multi': (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_vp, c_int, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp,
                                              c_vp, c_vp]),
    'ape_voxel_down_sample': (c_int, [c_vp, c_vp, c_int, c_dbl, c_vp, c_vp, c_vp]),
    'ape_voxel_down_sample_large': (c_int, [c_vp, c_int, c_dbl, c_vp, c_vp, c_vp]),
    'ape_pose_select': (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'ape_pose_compose': (c_int, [c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_int, c_vp, c_vp]),
    'ape_net_create': (c_int, [c_int, ctypes.POINTER(c_vp), c_int, c_int, c_int, c_int, ctypes.POINTER(c_vp)]),
    'ape_net_destroy': (c_int, [c_vp]),
    'ape_net_set_gemm': (c_int, [c_vp, c_int]),
    'ape_net_set_passes': (c_int, [c_vp, c_vp]),
    'ape_net_get_passes': (c_int, [c_vp, c_vp]),
    'ape_posenet_forward': (c_int, [c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'ape_refiner_forward': (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp]),
    'ape_pose_pipeline': (c_int, [c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int,
                                  c_vp, c_vp, c_vp]),
    'ape_posenet_forward_ex': (c_int, [c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'ape_pose_pipeline_ex': (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int,
                                     c_vp, c_vp, c_vp]),
    'ape_gather_emb': (c_int, [c_vp, c_int, c_int, c_vp, c_int, c_int, c_vp, c_vp]),
    'ape_host_gather_begin': (c_int, [c_vp, c_int, c_int, c_vp, c_int, c_int, c_int, c_vp, c_int]),
    'ape_host_gather_wait': (c_int, []),
    'ape_estimator_loss': (c_int, [c_vp] * 6 + [c_int, c_int, c_int, ctypes.c_float] + [c_vp] * 12),
    'ape_radius_outlier': (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_dbl, c_vp, c_vp, c_vp]),
    'ape_mahalanobis': (c_int, [c_vp, c_vp, c_int, c_vp, c_vp, c_vp]),
    'ape_statistical_outlier': (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_vp, c_dbl, c_vp, c_vp, c_vp, c_vp]),
    'ape_compact_points': (c_int, [c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp]),
    'ape_refiner_trainer_layout': (c_i64, [c_int, c_vp]),
    'ape_refiner_trainer_create': (c_int, [c_vp, c_vp, c_int, c_int, c_int, ctypes.POINTER(c_vp)]),
    'ape_refiner_trainer_destroy': (c_int, [c_vp]),
    'ape_refiner_trainer_sync_weights': (c_int, [c_vp, c_vp]),
    'ape_refiner_trainer_forward': (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp]),
    'ape_refiner_trainer_backward': (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp]),
    'ape_refiner_trainer_step': (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    'ape_refiner_trainer_adam': (c_int, [c_vp, c_vp, c_vp, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, c_int,
                                         ctypes.c_float, c_vp]),
    'ape_refiner_trainer_wait_bulk': (c_int, [c_vp, c_vp, c_vp]),
    'ape_refine_loss': (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'ape_adam_step': (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                              c_int, ctypes.c_float, c_vp]),
}

_lib = None


class ApeError(RuntimeError):
    pass


def load():
    """Load the shared library (building it is the job of __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ApeError('libape_b200.so is missing: run `python -m autoposeestimation_b200.build` '
                           '(there is no CPU fallback)')
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status, what=''):
    if status != 0:
        msg = load().ape_last_error()
        raise ApeError('%s failed (status %d): %s' % (what or 'ape call', status, msg.decode() if msg else ''))


def ptr(t):
    """Device (or host) pointer of a contiguous torch tensor; None -> NULL."""
    if t is None:
        return None
    assert t.is_contiguous(), 'tensor must be contiguous'
    return t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise ApeError('expected CUDA tensors: the B200 path has no CPU fallback')
