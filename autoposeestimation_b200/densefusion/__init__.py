"""Host-side mirror of the reference's DenseFusion call surface (lib/network.py, lib/knn, lib/loss*.py,
tools/utils.py) on top of the sm_100a kernels."""
from . import estimate_poses  # noqa: F401
