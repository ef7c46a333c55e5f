"""B200-native 6D-pose geometry hot path for KochPJ/AutoPoseEstimation (options 4 and 6).

Hand-written sm_100a CUDA kernels behind a C-ABI (include/ape_b200.h, libape_b200.so) and a
Python host layer that mirrors the reference's call surface.  No CPU fallback.
"""
from . import _lib  # noqa: F401

__all__ = ['_lib']
