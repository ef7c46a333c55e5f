"""Host-side mirror of pc_reconstruction/open3d_utils.py (get_surface, icp_regression) without open3d."""
