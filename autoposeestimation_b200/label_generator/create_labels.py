"""Drop-ins for the pose-label entry points of main.py option 4 ("Create Pose labels"):
`create_pose_label` (label_generator/create_labels.py:292-440) and `create_pose_data` (:40-290), same signatures.

The geometry -- point-to-point ICP of the aligned object cloud onto each rotation run's cloud, voxel grids -- runs on the
sm_100a kernels (pc_reconstruction.open3d_utils.icp_regression); the label algebra is a handful of 4x4 products per frame
on the host (formats.pose_label).  The smp U-Net of `create_pose_data` is outside the graft and is supplied by the caller.
Millimetres, fp64.  No CPU fallback."""
import json
import os

import numpy as np
import torch

from .. import formats
from ..pc_reconstruction.create_pointcloud import bbox_center, load_point_cloud
from ..pc_reconstruction.open3d_utils import PointCloud, icp_regression


def create_pose_label(root, object_name, global_regression, icp_point2point, icp_point2plane, plot=False, view_label=False,
                      with_extra=False):
    """create_labels.py:292-440.  For every rotation run of `data_generation/data/<object>/`: object position = bbox centre
    of the aligned cloud `<object>_out.ply`; for a run recorded with a non-zero object rotation the aligned cloud is
    registered onto that run's own cloud (`<run>.ply`, voxel 5, threshold 10) and the found rotation / centre refine the
    run's pose (only about the axes that were actually rotated, :363-372); then one `<id>.meta.json` pose label per frame
    (cam2object = inv(handEye) . inv(robot2endEff) . robot2object, :405-429).  `view_label` plots are outside the graft."""
    object_path = os.path.join(root, 'data_generation/data', object_name)
    runs = os.listdir(object_path)
    if 'background' not in runs:
        raise ValueError('background does not exist in object_path: {}'.format(object_path))
    runs.remove('background')
    if 'extra' in runs:
        runs.remove('extra')
        if with_extra:
            runs.append('extra')
    if len(runs) < 1:
        raise ValueError('no foreground')
    pc_dir = os.path.join(root, 'pc_reconstruction/data', object_name)
    aligned_path = os.path.join(pc_dir, '{}_out.ply'.format(object_name))
    remembered = []
    for d in runs:
        data_path = os.path.join(object_path, d)
        label_path = os.path.join(root, 'label_generator/data', object_name, d)
        pc_position = pc_rotation = None
        if d != 'extra':
            source = PointCloud(formats.read_ply(aligned_path))
            pc_position = bbox_center(source)
            for name in os.listdir(data_path):                 # the run's requested rotation: first meta file found (:336-341)
                if name.endswith('.json'):
                    with open(os.path.join(data_path, name)) as f:
                        pc_rotation = np.array(json.load(f).get('object_pose'), np.float64).reshape(4, 4)[:3, :3]
                    break
            requested = np.rad2deg(formats.mat2euler(pc_rotation))
            if not np.array_equal(requested, np.zeros(3)):
                target = PointCloud(formats.read_ply(os.path.join(pc_dir, '{}.ply'.format(d))))
                target, source, tf = icp_regression(target, source, voxel_size=5, threshold=10, global_regression=global_regression,
                                                    icp_point2point=icp_point2point, icp_point2plane=icp_point2plane, plot=False)
                euler = np.array(formats.mat2euler(np.dot(pc_rotation, tf[:3, :3])))
                euler[requested == 0.0] = 0.0                   # no correction about axes that were not rotated
                pc_rotation = formats.euler2mat(euler[0], euler[1], euler[2])
                pc_position = bbox_center(source)               # the down-sampled (not transformed) source, as :384
            remembered.append({'old_rotation': requested, 'pc_position': pc_position, 'pc_rotation': pc_rotation})
        ids = [f[:-10] for f in os.listdir(data_path) if '.color.png' in f]
        os.makedirs(label_path, exist_ok=True)
        for fid in ids:
            m = formats.load_frame_meta(os.path.join(data_path, '{}.meta.json'.format(fid)))
            if d == 'extra':
                rot = np.rad2deg(formats.mat2euler(m['object_rotation']))
                for r in remembered:
                    if np.array_equal(rot, r['old_rotation']):
                        pc_position, pc_rotation = r['pc_position'], r['pc_rotation']
                        break
            label = formats.pose_label(m['hand_eye'], m['robot2endEff'], pc_rotation, pc_position, object_name)
            formats.write_json(os.path.join(label_path, '{}.meta.json'.format(fid)), label)


def create_pose_data(root, classes, ds_name, reference_point=np.array([]), new_pred=True, get_extra_labels=False, plot=False,
                     use_cuda=True, segmentor_factory=None):
    """create_labels.py:40-290: per class (1) optionally refresh the segmentation labels with the trained U-Net
    (`<id>.new_pred.label.png`; largest connected component of the class, sanity checks against the background-subtraction
    label, the depth band around the reference point and the image centre, :96-213), (2) reconstruct the object cloud
    (`load_point_cloud` with the reference's fixed parameters, :217-253), (3) write the pose labels (`create_pose_label`).
    Step (1) is the reference's PyTorch / OpenCV code path and needs the smp model: pass
    `segmentor_factory(root, ds_name, n_classes)` (the reference's `get_default_model`); with new_pred=False and no
    'extra' run it is skipped and no segmentor is needed.  Returns the reference's statistics dict."""
    import cv2
    from torchvision import transforms
    from ..pipeline.utils import _largest_component_mask
    if not torch.cuda.is_available():
        raise RuntimeError('create_pose_data: a CUDA device is required (the B200 path has no CPU fallback)')
    device = torch.device('cuda:0')
    mode = 'new_pred' if new_pred else 'pred'
    to_tensor = transforms.ToTensor()
    normalize = transforms.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])
    model = None
    stats = {'n_samples': 0, 'n_extra_samples': 0, 'bs_copied': 0, 'no_depth_overlap': 0, 'not_in_center': 0}
    for class_id, cls in enumerate(classes):
        data_path = os.path.join(root, 'data_generation', 'data', cls)
        runs = [d for d in os.listdir(data_path) if d != 'background' and (get_extra_labels or d != 'extra')]
        for d in runs:
            if not (d == 'extra' or new_pred):
                continue
            if model is None:
                if segmentor_factory is None:
                    raise ValueError('create_pose_data: refreshing the segmentation labels needs segmentor_factory (smp U-Net, outside the graft)')
                model = segmentor_factory(root, ds_name, len(classes) + 1).to(device).eval()
            data_dir = os.path.join(data_path, d)
            label_path = os.path.join(root, 'label_generator/data', cls, d)
            os.makedirs(label_path, exist_ok=True)
            for fid in sorted(f[:-10] for f in os.listdir(data_dir) if '.color.png' in f):
                m = formats.load_frame_meta(os.path.join(data_dir, '{}.meta.json'.format(fid)))
                dist = np.linalg.norm(np.asarray(reference_point, np.float64) - m['robot2Cam'][:3, 3])
                depth = formats.load_depth_png(os.path.join(data_dir, '{}.depth.png'.format(fid))).astype(np.float64)
                depth[(depth > dist + 150) | (depth < dist - 150)] = 0
                rgb = cv2.cvtColor(cv2.imread(os.path.join(data_dir, '{}.color.png'.format(fid)), cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)
                with torch.no_grad():
                    prob = torch.softmax(model.predict(normalize(to_tensor(rgb)).to(device).unsqueeze(0)), dim=1)[0].cpu()
                arg = torch.argmax(prob, dim=0).numpy()
                arg = np.where(arg == class_id + 1, arg, 0)
                pred = _largest_component_mask(arg, arg * prob[class_id + 1].numpy())
                save = False
                if d != 'extra':
                    bs_label = formats.load_label_png(os.path.join(label_path, '{}.pred.label.png'.format(fid)))
                    if len(np.unique(pred[bs_label != 0])) <= 1:            # the network found nothing where the object is
                        pred, save = bs_label, True
                        stats['bs_copied'] += 1
                if not save:
                    if len(np.unique(pred[depth != 0])) <= 1:
                        stats['no_depth_overlap'] += 1
                    elif len(np.unique(pred[30:pred.shape[0] - 30, 50:pred.shape[1] - 50])) > 1:
                        save = True
                    else:
                        stats['not_in_center'] += 1
                new_label = os.path.join(label_path, '{}.new_pred.label.png'.format(fid))
                if save:
                    stats['n_extra_samples' if d == 'extra' else 'n_samples'] += 1
                    formats.save_png(new_label, pred)
                else:
                    for stale in (new_label, os.path.join(label_path, '{}.meta.json'.format(fid))):
                        if os.path.exists(stale):
                            os.remove(stale)
        load_point_cloud(cls, os.path.join(root, 'pc_reconstruction/data'), root, reference_point=reference_point, mode=mode,
                         n_viewpoints=30, min_friends=20, min_dist=5, nb_neighbors=20, threshold=10, voxel_size=2, voxel_size_out=5,
                         l_arrow=75, global_regression=False, icp_point2point=True, icp_point2plane=False, plot=False)
        create_pose_label(root, cls, False, True, False, plot=False, view_label=False, with_extra=get_extra_labels)
    return stats
