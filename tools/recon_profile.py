"""Per-kernel device time of the 30-view register-and-merge loop (ape_reconstruct_run): python tools/recon_profile.py"""
import sys, time
sys.path.insert(0, '/root/repo' if __name__ == '__main__' else '.')
import torch
import bench
from autoposeestimation_b200 import _lib, synthetic as synth
from autoposeestimation_b200.pc_reconstruction.create_pointcloud import get_surfaces_batch, reconstruct_run
lib = _lib.load()
dev = torch.device('cuda', 0)
n_views = 30
scene = synth.Scene(5, n_objects=1)
poses = scene.camera_poses(21, n_views)
lab, dep = scene.render(poses, seed=3, device=dev, only_object=0)
cam = torch.tensor([[synth.INTR['ppx'], synth.INTR['ppy'], synth.INTR['fx'], synth.INTR['fy']]], dtype=torch.float64, device=dev).repeat(n_views, 1)
r2c = torch.from_numpy(poses).to(dev)
surfaces = get_surfaces_batch(lab, dep, cam, r2c, 20, 5.0, 20, 2.0)
reconstruct_run(surfaces, 2.0, 10.0)
torch.cuda.synchronize()
lib.ape_profile_enable(1)
reconstruct_run(surfaces, 2.0, 10.0)
torch.cuda.synchronize()
rep = bench.profile_report(lib); lib.ape_profile_enable(0)
for k, (n, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
    print('%-28s %4d launches %8.3f ms total %8.1f us each' % (k, n, ms, ms / n * 1e3))
# iterations per registration (view by view, for the record)
from autoposeestimation_b200.pc_reconstruction.open3d_utils import icp_regression_batch, PointCloud
live = [s for s in surfaces if len(s) > 0]
cloud = PointCloud(live[0].points.clone()); its = []
for s_ in live[1:]:
    td, sd, T, info = icp_regression_batch([cloud], [s_], 2.0, 10.0)
    its.append((int(info[0, 2]), len(sd[0]), len(td[0])))
    cloud = PointCloud(torch.cat((sd[0].transform(T[0]).points, td[0].points))).voxel_down_sample(2.0)
print('iterations, source points, target points per view:', its)
