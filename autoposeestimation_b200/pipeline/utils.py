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
    out = {}
    # crops differ in size per object, so the encoder (PyTorch/cuDNN, out of the graft) runs per crop;
    # the geometry kernels then run once per distinct crop size
    groups = {}
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


def _run_groups(image_chw, bboxes, cloud, choose_t, idx, sel, estimator, refiner, num_points, refine_mode, iterations):
    """Encoder per distinct crop size (PyTorch/cuDNN, outside the graft), geometry kernels once per group."""
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
