"""Option 6 ("Run Live Prediction") geometry block: mask -> bbox -> choose -> back-projection -> PoseNet ->
refinement -> pose, for all detected objects of a frame in one batched pass.

Reference: pipeline/utils.py:517-574 inside full_prediction (the segmentation / connected-components part
before it, :422-469, stays on PyTorch/OpenCV and is out of scope).  Differences from the reference, all
deliberate and documented in DESIGN.md:
  * xmap/ymap are never built (the reference rebuilds two 480x640 index maps with Python list comprehensions
    per call, :518-519); the kernel derives row/col from `choose`;
  * all objects of the frame go through ONE batched launch sequence and ONE D2H copy instead of >= 3 H2D and
    >= 3 D2H syncs per object;
  * `refine_mode='live'` reproduces :569-571 exactly (two identical refiner calls -> one composition);
    `refine_mode='canonical'` follows DenseFusion/tools/eval_linemod.py:91-114.
"""
import numpy as np
import torch

from .. import ops

IMG_H, IMG_W, BORDER = 480, 640, 40


def get_bbox(label):
    """DenseFusion/datasets/myDatasetAugmented/dataset.py:342-380 (border_list :338): tight box, each side
    length raised to the next multiple of 40 unless it already is one, re-centred, shifted back inside 480x640."""
    label = np.asarray(label) != 0
    rr = np.flatnonzero(label.any(axis=1)); cc = np.flatnonzero(label.any(axis=0))
    box = []
    for lo, hi, limit in ((int(rr[0]), int(rr[-1]) + 1, IMG_H), (int(cc[0]), int(cc[-1]) + 1, IMG_W)):
        ext = hi - lo
        if ext % BORDER:
            ext = (ext // BORDER + 1) * BORDER
        mid = (lo + hi) // 2
        lo, hi = mid - ext // 2, mid + ext // 2
        if lo < 0:
            lo, hi = 0, hi - lo
        if hi > limit:
            lo, hi = lo - (hi - limit), limit
        box += [lo, hi]
    return tuple(box)


def choose_points(mask, bbox, num_points, rng=np.random):
    """pipeline/utils.py:529-539: flat indices of mask pixels inside the crop; more than num_points -> keep a
    random subset in ascending order (np.random.shuffle of a 0/1 vector); fewer -> cyclic 'wrap' padding."""
    rmin, rmax, cmin, cmax = bbox
    choose = np.flatnonzero(mask[rmin:rmax, cmin:cmax].ravel())
    if len(choose) == 0:
        return None
    if len(choose) > num_points:
        keep = np.zeros(len(choose), dtype=int)
        keep[:num_points] = 1
        rng.shuffle(keep)
        return choose[keep.nonzero()]
    return np.pad(choose, (0, num_points - len(choose)), 'wrap')


def predict_poses(image_chw, depth, meta, masks, class_ids, estimator, refiner, num_points=1000, refine_mode='live',
                  iterations=2, rng=np.random, device_sampling=False, seed=0):
    """Geometry block of full_prediction for the objects of one frame.

    image_chw : normalised colour image as a [3,480,640] float CUDA tensor (what `normalize(to_tensor)` gives)
    depth     : [480,640] uint16 numpy;  meta: {'intr': {ppx,ppy,fx,fy}, 'depth_scale': float}
    masks     : list of [480,640] uint8 numpy (255 = object) ; class_ids: list of int
    estimator / refiner : densefusion.network.PoseNet / PoseRefineNet (CUDA, eval mode)
    device_sampling : mask -> bbox -> choose -> back-projection run in ONE kernel on the device (ops.mask_bbox_choose,
                SURVEY 8f rank 2); the random subset then comes from a seeded hash instead of numpy's global RNG (same
                distribution).  masks may then also be a [n,480,640] uint8 CUDA tensor (e.g. straight from the segmentor).
    Returns {i: {'position': np[3] (m), 'rotation': np[4] wxyz}} for every object with at least one valid pixel
    (objects without one are skipped, as :530-531)."""
    dev = image_chw.device
    if device_sampling:
        return _predict_poses_device(image_chw, depth, meta, masks, class_ids, estimator, refiner, num_points, refine_mode,
                                     iterations, seed)
    depth = np.ascontiguousarray(depth)
    sel, bboxes, chooses = [], [], []
    for i, m in enumerate(masks):
        mask_label = np.asarray(m) == 255
        if not mask_label.any():
            continue
        bbox = get_bbox(mask_label)
        ch = choose_points(mask_label & (depth != 0), bbox, num_points, rng)
        if ch is None:
            continue
        sel.append(i); bboxes.append(bbox); chooses.append(ch)
    if not sel:
        return {}
    B = len(sel)
    intr = meta['intr']
    cam = np.tile(np.array([[intr['ppx'], intr['ppy'], intr['fx'], intr['fy'], meta['depth_scale']]], np.float32), (B, 1))
    d16 = torch.from_numpy(depth.astype(np.uint16).view(np.int16)[None].copy()).to(dev)
    choose_t = torch.from_numpy(np.stack(chooses).astype(np.int64)).to(dev)
    cloud = ops.backproject_choose(d16, torch.tensor(bboxes, dtype=torch.int32, device=dev), choose_t,
                                   torch.from_numpy(cam).to(dev), frame_of=torch.zeros(B, dtype=torch.int32, device=dev))
    idx = torch.tensor([class_ids[i] for i in sel], dtype=torch.int64, device=dev)
    return _run_groups(image_chw, bboxes, cloud, choose_t, idx, sel, estimator, refiner, num_points, refine_mode, iterations)


def _run_groups(image_chw, bboxes, cloud, choose_t, idx, sel, estimator, refiner, num_points, refine_mode, iterations):
    """Crops differ in size per object, so the colour encoder (PyTorch/cuDNN, outside the graft) runs once per distinct crop
    size; the geometry kernels then run once per group and ONE D2H copy returns the group's poses."""
    dev = image_chw.device
    out, groups = {}, {}
    for j, bb in enumerate(bboxes):
        groups.setdefault((bb[1] - bb[0], bb[3] - bb[2]), []).append(j)
    with torch.no_grad():
        for (h, w), js in groups.items():
            crops = torch.stack([image_chw[:, bboxes[j][0]:bboxes[j][1], bboxes[j][2]:bboxes[j][3]] for j in js])
            out_img = estimator.cnn(crops)
            jt = torch.tensor(js, device=dev)
            est_h = estimator._handle(len(js), num_points)
            ref_h = refiner._handle(len(js), num_points) if refiner is not None and iterations > 0 else None
            poses, _ = ops.pose_pipeline(est_h, ref_h, out_img, cloud[jt], choose_t[jt], idx[jt],
                                         iterations=iterations if ref_h is not None else 0, canonical=(refine_mode == 'canonical'))
            poses = poses.cpu().numpy()
            for k, j in enumerate(js):
                out[sel[j]] = {'rotation': poses[k, :4].copy(), 'position': poses[k, 4:].copy()}
    return out


def _predict_poses_device(image_chw, depth, meta, masks, class_ids, estimator, refiner, num_points, refine_mode, iterations, seed):
    dev = image_chw.device
    if isinstance(masks, torch.Tensor):
        lab = masks.to(dev, torch.uint8).contiguous()
    else:
        if len(masks) == 0:
            return {}
        lab = torch.from_numpy(np.ascontiguousarray(np.stack([np.asarray(m, np.uint8) for m in masks]))).to(dev)
    n = lab.shape[0]
    if n == 0:
        return {}
    if isinstance(depth, torch.Tensor):
        d16 = depth.to(dev).reshape(1, IMG_H, IMG_W).contiguous()
    else:
        d16 = torch.from_numpy(np.ascontiguousarray(depth).astype(np.uint16).view(np.int16)[None].copy()).to(dev)
    intr = meta['intr']
    cam = torch.tensor([[intr['ppx'], intr['ppy'], intr['fx'], intr['fy'], meta['depth_scale']]], dtype=torch.float32, device=dev).repeat(n, 1)
    # object i = mask i (its own label plane) over the one depth frame
    depth_n = d16.expand(n, IMG_H, IMG_W).contiguous() if n > 1 else d16
    seeds = torch.arange(n, device=dev, dtype=torch.int64) * 0x9E3779B1 + int(seed)
    r = ops.mask_bbox_choose(lab, depth_n, cam, num_points, seeds=seeds)
    ncand = r['n_candidates'].cpu().numpy()                 # the one host sync: which objects exist + their crop boxes
    bbox = r['bbox'].cpu().numpy()
    sel = [i for i in range(n) if ncand[i] > 0]
    if not sel:
        return {}
    st = torch.tensor(sel, device=dev)
    idx = torch.tensor([class_ids[i] for i in sel], dtype=torch.int64, device=dev)
    return _run_groups(image_chw, [tuple(int(v) for v in bbox[i]) for i in sel], r['cloud'][st], r['choose'][st], idx, sel,
                       estimator, refiner, num_points, refine_mode, iterations)


# ------------------------------------------------------------------------------------------------------------------
# Entry points of main.py option 6 with the reference's signatures (pipeline/utils.py:410-641, :643-718)
def _largest_component_mask(cls_arg, cls_score):
    """pipeline/utils.py:447-466: 8-connected components of the class's arg-max mask (cv2.connectedComponents, as the
    reference), keep the component with the highest MEAN class probability, return it as a 0 / 255 uint8 mask."""
    import cv2
    _, labels = cv2.connectedComponents(cls_arg.astype(np.uint8), connectivity=8)
    best, best_score = 1, 0.0
    for u in np.unique(labels):
        if u == 0:
            continue
        score = float(np.mean(cls_score[labels == u]))
        if score > best_score:
            best_score, best = score, u
    return np.where((labels == best) & (cls_score != 0), 255, 0).astype(np.uint8)


def segment_objects(image, segmentor, to_tensor, normalize, device, class_names, min_pixels=100):
    """Segmentation front half of full_prediction (:425-469; smp U-Net on PyTorch/cuDNN + OpenCV, outside the graft):
    -> ({class name: 0/255 mask}, per-pixel arg-max) in ascending class-id order."""
    x = normalize(to_tensor(image.copy())).to(device).unsqueeze(0)
    with torch.no_grad():
        prob = torch.softmax(segmentor.predict(x), dim=1)[0].cpu()
    arg = torch.argmax(prob, dim=0).numpy()
    masks = {}
    found, counts = np.unique(arg, return_counts=True)
    for cls, cnt in zip(found, counts):
        if cls == 0 or cnt <= min_pixels:
            continue
        cls_arg = np.where(arg == cls, arg, 0)
        masks[class_names[cls - 1]] = _largest_component_mask(cls_arg, cls_arg * prob[cls].numpy())
    return masks, arg


def project_cloud_to_image(image, points, radius, intr, color=(255, 0, 0)):
    """`pc_utils.pointcloud2image` as full_prediction uses it (:583-587): draw the posed model cloud (metres, camera frame)
    into the image with the pinhole model; host-side visualisation only."""
    import cv2
    pts = np.asarray(points, np.float64)
    pts = pts[pts[:, 2] > 0]
    u = np.round(pts[:, 0] / pts[:, 2] * intr['fx'] + intr['ppx']).astype(int)
    v = np.round(pts[:, 1] / pts[:, 2] * intr['fy'] + intr['ppy']).astype(int)
    ok = (u >= 0) & (u < image.shape[1]) & (v >= 0) & (v < image.shape[0])
    canvas = np.ascontiguousarray(image)
    for x, y in zip(u[ok], v[ok]):
        cv2.circle(canvas, (int(x), int(y)), int(radius), tuple(float(c) for c in color), -1)
    return canvas


def full_prediction(image, depth, meta, segmentor, estimator, refiner, to_tensor, normalize, device, cuda, color_dict,
                    class_names=None, point_clouds=None, plot=False, color_prediction=False, bbox=False, put_text=False):
    """Drop-in for pipeline/utils.py:410-641 (main.py option 6, grasping_utils.py:80): same arguments, same result dict
    {'predictions': {cls: {'mask' u8 [480,640], 'position' np[3] (m), 'rotation' np[4] wxyz}}, 'elapsed_times': {...}
    [, 'segmented_prediction', 'pose_prediction']}.

    The segmentation half (U-Net, soft-max, connected components) stays on PyTorch / OpenCV as in the reference; the geometry
    half (:517-574: mask -> bbox -> choose -> back-projection -> PoseNet -> refinement) is ONE batched pass over all detected
    objects on the sm_100a kernels (`predict_poses`, refine_mode='live' = the reference's loop as written), with ONE
    device-to-host copy instead of >= 3 syncs per object.  `choose` consumes numpy's global RNG in the reference's order
    (objects in ascending class id, np.random.shuffle only for objects with more than 1000 candidate pixels), so a seeded
    run selects the same pixels.  `cuda` is accepted for signature compatibility: there is no CPU path (raises without a GPU).
    matplotlib plots (plot=True, :596-611) are outside the graft and ignored."""
    import time
    import cv2
    t0 = time.time()
    if not torch.cuda.is_available() or torch.device(device).type != 'cuda':
        raise ops._lib.ApeError('full_prediction: a CUDA device is required (the B200 path has no CPU fallback)')
    out = {'predictions': {}, 'elapsed_times': {}}
    img_np = np.array(image)
    if color_prediction:
        out['segmented_prediction'] = img_np.astype(np.float64).copy()
        out['pose_prediction'] = img_np.astype(np.float64).copy()
    masks, _ = segment_objects(img_np, segmentor, to_tensor, normalize, device, class_names)
    for cls, m in masks.items():
        out['predictions'][cls] = {'mask': m}
        if color_prediction:
            col = color_dict[cls]['value']
            sel = m != 0
            for c, v in enumerate(col):
                out['segmented_prediction'][:, :, c][sel] = out['segmented_prediction'][:, :, c][sel] * 0.7 + v * 0.3
            if bbox or put_text:
                bb = get_bbox(m)
                if bbox:
                    cv2.rectangle(out['segmented_prediction'], (bb[2], bb[0]), (bb[3], bb[1]), col, 2)
                if put_text:
                    cv2.putText(out['segmented_prediction'], 'Segmentation', (10, 30), cv2.FONT_HERSHEY_SIMPLEX, 1, (0, 0, 0), 2, cv2.LINE_AA)
                    cv2.putText(out['segmented_prediction'], cls, (bb[2] + 10, bb[0] - 10), cv2.FONT_HERSHEY_SIMPLEX, 1, col, 2, cv2.LINE_AA)
    if color_prediction:
        out['segmented_prediction'] = np.clip(out['segmented_prediction'], 0, 255).astype(np.uint8)
    out['elapsed_times']['segmentation'] = time.time() - t0

    t1 = time.time()
    names = list(out['predictions'])
    if names:
        # the estimator sees the RAW 0..255 crop passed through `normalize` only (:559-560: no to_tensor scaling)
        raw = torch.from_numpy(np.ascontiguousarray(np.transpose(img_np[:, :, :3], (2, 0, 1))).astype(np.float32))
        image_chw = normalize(raw).to(device)
        poses = predict_poses(image_chw, np.asarray(depth), meta, [out['predictions'][c]['mask'] for c in names],
                              [int(class_names.index(c)) for c in names], estimator, refiner, num_points=1000,
                              refine_mode='live', iterations=2, rng=np.random)
        for j, cls in enumerate(names):
            if j in poses:
                out['predictions'][cls]['position'] = poses[j]['position']
                out['predictions'][cls]['rotation'] = poses[j]['rotation']
                if color_prediction and point_clouds is not None:
                    from ..densefusion.transformations import quaternion_matrix
                    R = quaternion_matrix(poses[j]['rotation'])[:3, :3]
                    posed = np.dot(point_clouds[class_names.index(cls)], R.T) + poses[j]['position']
                    out['pose_prediction'] = project_cloud_to_image(out['pose_prediction'], posed, 3, meta['intr'], color=color_dict[cls]['value'])
                    if put_text:
                        cv2.putText(out['pose_prediction'], 'Pose Estimation', (10, 30), cv2.FONT_HERSHEY_SIMPLEX, 1, (0, 0, 0), 2, cv2.LINE_AA)
    if color_prediction:
        out['pose_prediction'] = np.clip(out['pose_prediction'], 0, 255).astype(np.uint8)
    out['elapsed_times']['pose_estimation'] = time.time() - t1
    for cls in [c for c, v in out['predictions'].items() if 'position' not in v or 'rotation' not in v]:
        print('Deleting cls "{}"'.format(cls))                # as :624-626 (objects without a valid depth pixel)
        del out['predictions'][cls]
    out['elapsed_times']['total'] = time.time() - t0
    return out


def get_prediction_models(root, data_set_name, segmentor_factory=None):
    """Drop-in for pipeline/utils.py:643-718: -> (segmentor, estimator, refiner, classes, to_tensor, normalize, cld,
    device, cuda).  Reads `label_generator/data_sets/segmentation/<ds>/classes.txt`, the `.xyz` model cloud of every class
    (metres, parsed exactly as :667-684 incl. its dropped-character quirk), and the DenseFusion checkpoints
    `DenseFusion/trained_models/<ds>/pose_model.pth` / `pose_refine_model.pth` into the grafted PoseNet / PoseRefineNet
    (num_points 1000).  The smp U-Net segmentor is outside the graft: pass `segmentor_factory(root, data_set_name,
    n_classes)` (the reference's `segmentation.utils.get_default_model`); without one the first tuple element is None."""
    import os
    from torchvision import transforms
    from .. import formats
    from ..densefusion.network import PoseNet, PoseRefineNet
    if not torch.cuda.is_available():
        raise ops._lib.ApeError('get_prediction_models: a CUDA device is required (the B200 path has no CPU fallback)')
    device, cuda = torch.device('cuda:0'), True
    classes, cld = [], {}
    with open(os.path.join(root, 'label_generator', 'data_sets', 'segmentation', data_set_name, 'classes.txt')) as f:
        for line in f:
            name = line.rstrip('\n')
            if not name:
                break
            cld[len(classes)] = formats.read_xyz(os.path.join(root, 'pc_reconstruction', 'data', name, '{}.xyz'.format(name)), to_meter=True)
            classes.append(name)
    to_tensor = transforms.ToTensor()
    normalize = transforms.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])
    segmentor = None
    if segmentor_factory is not None:
        segmentor = segmentor_factory(root, data_set_name, len(classes) + 1).to(device).eval()
    pose_path = os.path.join(root, 'DenseFusion', 'trained_models', data_set_name)
    estimator = PoseNet(num_points=1000, num_obj=len(classes))
    refiner = PoseRefineNet(num_points=1000, num_obj=len(classes))
    estimator.load_state_dict(torch.load(os.path.join(pose_path, 'pose_model.pth'), map_location='cpu'))
    refiner.load_state_dict(torch.load(os.path.join(pose_path, 'pose_refine_model.pth'), map_location='cpu'))
    estimator.to(device).eval(); refiner.to(device).eval()
    return segmentor, estimator, refiner, classes, to_tensor, normalize, cld, device, cuda
