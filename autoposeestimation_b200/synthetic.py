"""Seeded synthetic inputs for the BASELINE.json configs (test and bench infrastructure; no oracle code).

Everything is generated with numpy's frozen legacy ``RandomState`` streams so the
same seed gives bit-identical inputs in the dev container and on the GPU box
(the golden fixtures under tests/golden store only seeds + reference outputs).
Shapes follow SURVEY.md section 8(d).
"""
import math
import numpy as np

# RealSense-like intrinsics used by every config
INTR = dict(ppx=320.0, ppy=240.0, fx=615.0, fy=615.0)
DEPTH_SCALE = 0.001                     # metres per raw unit (DenseFusion path works in metres)

# hand_eye_calibration/data/handEye_tf.json-like extrinsic (mm); values are synthetic but
# of the same magnitude (translation about -88/-31/-188 mm)
def hand_eye():
    a = math.radians(1.5)
    T = np.identity(4)
    T[:3, :3] = np.array([[math.cos(a), -math.sin(a), 0.0], [math.sin(a), math.cos(a), 0.0], [0.0, 0.0, 1.0]])
    T[:3, 3] = [-88.0, -31.0, -188.0]
    return T


_POSENET_SHAPES = [
    ('feat.conv1', 64, 3), ('feat.conv2', 128, 64), ('feat.e_conv1', 64, 32), ('feat.e_conv2', 128, 64),
    ('feat.conv5', 512, 256), ('feat.conv6', 1024, 512),
    ('conv1_r', 640, 1408), ('conv1_t', 640, 1408), ('conv1_c', 640, 1408),
    ('conv2_r', 256, 640), ('conv2_t', 256, 640), ('conv2_c', 256, 640),
    ('conv3_r', 128, 256), ('conv3_t', 128, 256), ('conv3_c', 128, 256),
]
_REFINER_SHAPES = [
    ('feat.conv1', 64, 3), ('feat.conv2', 128, 64), ('feat.e_conv1', 64, 32), ('feat.e_conv2', 128, 64),
    ('feat.conv5', 512, 384), ('feat.conv6', 1024, 512),
    ('conv1_r', 512, 1024), ('conv1_t', 512, 1024), ('conv2_r', 128, 512), ('conv2_t', 128, 512),
]


def _uniform_layer(rng, cout, cin, conv):
    b = 1.0 / math.sqrt(cin)            # torch's default Conv1d/Linear init bound
    w = rng.uniform(-b, b, size=(cout, cin)).astype(np.float32)
    bias = rng.uniform(-b, b, size=(cout,)).astype(np.float32)
    return (w[:, :, None] if conv else w), bias


def posenet_state_dict(seed, num_obj):
    """Reference-shaped state_dict (numpy fp32) of PoseNet *without* the colour encoder
    (network.py:74-91).  Conv1d weights are [Cout,Cin,1]."""
    rng = np.random.RandomState(seed)
    sd = {}
    for name, co, ci in _POSENET_SHAPES:
        sd[name + '.weight'], sd[name + '.bias'] = _uniform_layer(rng, co, ci, True)
    for h, wd in (('r', 4), ('t', 3), ('c', 1)):
        sd['conv4_%s.weight' % h], sd['conv4_%s.bias' % h] = _uniform_layer(rng, num_obj * wd, 128, True)
    return sd


def refiner_state_dict(seed, num_obj):
    """Reference-shaped state_dict of PoseRefineNet (network.py:139-183)."""
    rng = np.random.RandomState(seed)
    sd = {}
    for name, co, ci in _REFINER_SHAPES:
        sd[name + '.weight'], sd[name + '.bias'] = _uniform_layer(rng, co, ci, name.startswith('feat.'))
    for h, wd in (('r', 4), ('t', 3)):
        sd['conv3_%s.weight' % h], sd['conv3_%s.bias' % h] = _uniform_layer(rng, num_obj * wd, 128, False)
    return sd


def to_torch(sd):
    import torch
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}


def posenet_inputs(seed, n_points, crop_hw=(120, 160), num_obj=5, batch=1):
    """encoder output ~N(0,1) [B,32,Hc,Wc] fp32, choose = sorted random n of
    Hc*Wc, depth-consistent cloud [B,n,3] (z in [0.4,0.8] m), idx uniform in [0,num_obj)."""
    rng = np.random.RandomState(seed)
    hc, wc = crop_hw
    out_img = rng.standard_normal((batch, 32, hc, wc)).astype(np.float32)
    choose = np.stack([np.sort(rng.choice(hc * wc, n_points, replace=False)) for _ in range(batch)]).astype(np.int64)
    z = rng.uniform(0.4, 0.8, size=(batch, 1)).astype(np.float32) + \
        rng.uniform(-0.03, 0.03, size=(batch, n_points)).astype(np.float32)
    rows = (choose // wc).astype(np.float32) + 180.0
    cols = (choose % wc).astype(np.float32) + 240.0
    x = (cols - np.float32(INTR['ppx'])) * z / np.float32(INTR['fx'])
    y = (rows - np.float32(INTR['ppy'])) * z / np.float32(INTR['fy'])
    cloud = np.stack([x, y, z], axis=2).astype(np.float32)
    idx = rng.randint(0, num_obj, size=(batch, 1)).astype(np.int64)
    return out_img, cloud, choose[:, None, :], idx


def random_rotation(rng, max_angle_rad):
    axis = rng.standard_normal(3); axis /= np.linalg.norm(axis)
    ang = rng.uniform(-max_angle_rad, max_angle_rad)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.identity(3) + math.sin(ang) * K + (1 - math.cos(ang)) * (K @ K)


def ellipsoid_cloud(rng, n, semi_axes=(40.0, 30.0, 20.0)):
    """n points uniform-ish on an ellipsoid surface (mm)."""
    v = rng.standard_normal((n, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
    return v * np.asarray(semi_axes)


def render_ellipsoid_frame(seed, semi_axes=(40.0, 30.0, 20.0), H=480, W=640, noise_mm=0.5, dropout=0.02,
                           max_rot_deg=10.0, max_trans_mm=5.0, n_model=2000):
    """one synthetic 640x480 depth frame (u16, mm) of an ellipsoid at 500-700 mm,
    label u8 (255 on the object), the 2000-pt model cloud in the robot frame, the extrinsic,
    and the ground-truth perturbation that ICP has to recover.  Ray-casting is analytic."""
    rng = np.random.RandomState(seed)
    a = np.asarray(semi_axes, np.float64)
    cam_c = np.array([rng.uniform(-60, 60), rng.uniform(-40, 40), rng.uniform(500, 700)])   # centre, camera frame
    R_obj = random_rotation(rng, math.pi)
    fx, fy, ppx, ppy = INTR['fx'], INTR['fy'], INTR['ppx'], INTR['ppy']
    vv, uu = np.mgrid[0:H, 0:W]
    d = np.stack([(uu - ppx) / fx, (vv - ppy) / fy, np.ones_like(uu, dtype=np.float64)], axis=-1)   # z-normalised rays
    # ellipsoid: |A R^T (s d - c)|^2 = 1, A = diag(1/a)
    M = (R_obj / a).T                      # rows: (R[:,i]/a_i)^T
    dd = d @ M.T
    cc = M @ cam_c
    qa = (dd * dd).sum(-1); qb = -2.0 * (dd @ cc); qc = float(cc @ cc) - 1.0
    disc = qb * qb - 4 * qa * qc
    hit = disc > 0
    s = np.where(hit, (-qb - np.sqrt(np.where(hit, disc, 0.0))) / (2 * qa), 0.0)   # z depth of first hit
    z = s + rng.standard_normal(s.shape) * noise_mm
    depth = np.zeros((H, W), np.uint16)
    bg = rng.uniform(900, 1100, size=(H, W))
    depth[:] = np.round(bg).astype(np.uint16)
    depth[hit] = np.clip(np.round(z[hit]), 1, 65535).astype(np.uint16)
    depth[rng.uniform(size=(H, W)) < dropout] = 0
    label = np.where(hit, 255, 0).astype(np.uint8)
    robot2cam = hand_eye()
    # model cloud in robot frame = object surface mapped through the extrinsic, then perturbed
    model_obj = ellipsoid_cloud(rng, n_model, semi_axes)
    model_cam = model_obj @ R_obj.T + cam_c
    model_robot = model_cam @ robot2cam[:3, :3].T + robot2cam[:3, 3]
    dR = random_rotation(rng, math.radians(max_rot_deg))
    dt = rng.uniform(-max_trans_mm, max_trans_mm, size=3) / math.sqrt(3.0)
    ctr = model_robot.mean(axis=0)
    model_pert = (model_robot - ctr) @ dR.T + ctr + dt
    return dict(depth=depth, label=label, intr=dict(INTR), robot2cam=robot2cam,
                model=model_pert, model_true=model_robot, dR=dR, dt=dt)


def adds_instances(seed, n_inst, n_models=21, n_model_pts=2600, n_pred_pts=500, n_sym=5):
    """shared model clouds [n_models, 2600, 3] fp32 in a 0.2 m cube; per instance
    class id, GT pose, predicted pose = GT o small perturbation, 500-pt subsample indices."""
    rng = np.random.RandomState(seed)
    models = rng.uniform(-0.1, 0.1, size=(n_models, n_model_pts, 3)).astype(np.float32)
    cls = rng.randint(0, n_models, size=n_inst).astype(np.int32)
    def quats(maxang):
        ax = rng.standard_normal((n_inst, 3)); ax /= np.linalg.norm(ax, axis=1, keepdims=True)
        ang = rng.uniform(-maxang, maxang, size=(n_inst, 1))
        return np.concatenate([np.cos(ang / 2), ax * np.sin(ang / 2)], axis=1)
    q_gt = quats(math.pi)
    t_gt = rng.uniform(-0.3, 0.3, size=(n_inst, 3)); t_gt[:, 2] += 0.7
    dq = quats(math.radians(5.0))
    w1, v1 = q_gt[:, :1], q_gt[:, 1:]; w2, v2 = dq[:, :1], dq[:, 1:]
    q_pr = np.concatenate([w1 * w2 - (v1 * v2).sum(1, keepdims=True), w1 * v2 + w2 * v1 + np.cross(v1, v2)], axis=1)
    t_pr = t_gt + rng.uniform(-0.01, 0.01, size=(n_inst, 3)) / math.sqrt(3.0)
    sub = np.sort(rng.choice(n_model_pts, n_pred_pts, replace=False)).astype(np.int32)
    sym = np.zeros(n_models, np.uint8); sym[:n_sym] = 1
    return dict(models=models, cls=cls, q_gt=q_gt.astype(np.float32), t_gt=t_gt.astype(np.float32),
                q_pred=q_pr.astype(np.float32), t_pred=t_pr.astype(np.float32), subsample=sub, sym=sym)


def encoder_state_dict(seed, shapes):
    """Seeded weights for the colour encoder (keys/shapes as given, e.g. from tests/golden/state_dict_shapes.json).
    Scaled like He-init so activations stay O(1) through the 17 conv layers."""
    rng = np.random.RandomState(seed)
    sd = {}
    for k in sorted(shapes):
        shp = tuple(shapes[k])
        if k.endswith('.bias'):
            sd[k] = (rng.standard_normal(shp) * 0.01).astype(np.float32)
        elif len(shp) == 1:                      # PReLU slope
            sd[k] = np.full(shp, 0.25, np.float32)
        else:
            fan_in = int(np.prod(shp[1:]))
            sd[k] = (rng.standard_normal(shp) * math.sqrt(2.0 / fan_in)).astype(np.float32)
    return sd


# ------------------------------------------------------------------------------------------------------------------
# Multi-object scene seen from many camera poses ("10 000 random camera poses on a hemisphere r in
# [450,700] mm looking at the table centre; each frame contains all 5 objects' labels")
SCENE_AXES = ((40.0, 30.0, 20.0), (35.0, 35.0, 25.0), (50.0, 25.0, 20.0), (30.0, 30.0, 28.0), (45.0, 22.0, 26.0))
SCENE_CENTER = np.array([0.0, -500.0, 60.0])                  # table centre in the robot frame (mm)


class Scene:
    """n_objects ellipsoids with distinct semi-axes at fixed poses around SCENE_CENTER (robot frame, mm).
    `object_rotation` (optional 3x3) rotates every object about its own centre (a "rotation run" of the turntable)."""

    def __init__(self, seed=3, n_objects=5, n_model=2000, object_rotation=None):
        rng = np.random.RandomState(seed)
        self.n = n_objects
        self.axes = np.array(SCENE_AXES[:n_objects])
        ang = np.arange(n_objects) * (2 * math.pi / max(n_objects, 1)) + rng.uniform(0, 1)
        radius = 0.0 if n_objects == 1 else 115.0
        self.centers = SCENE_CENTER + np.stack([radius * np.cos(ang), radius * np.sin(ang), np.zeros(n_objects)], 1)
        self.R = np.stack([random_rotation(rng, math.pi) for _ in range(n_objects)])
        if object_rotation is not None:
            self.R = np.stack([np.asarray(object_rotation) @ r for r in self.R])
        # model clouds (true surface, robot frame) and perturbed copies (<= 10 deg, <= 5 mm) that ICP has to bring back
        self.models, self.models_pert = [], []
        for k in range(n_objects):
            m = ellipsoid_cloud(rng, n_model, self.axes[k]) @ self.R[k].T + self.centers[k]
            dR = random_rotation(rng, math.radians(10.0)); dt = rng.uniform(-5.0, 5.0, size=3) / math.sqrt(3.0)
            self.models.append(m)
            self.models_pert.append((m - self.centers[k]) @ dR.T + self.centers[k] + dt)

    def camera_poses(self, seed, n):
        """n camera->robot transforms [n,4,4]: positions on the upper hemisphere around SCENE_CENTER (r in [450,700] mm,
        elevation 35..80 deg), optical axis through the centre, small in-plane roll."""
        rng = np.random.RandomState(seed)
        r = rng.uniform(450.0, 700.0, n); el = np.radians(rng.uniform(35.0, 80.0, n)); az = rng.uniform(0, 2 * math.pi, n)
        roll = np.radians(rng.uniform(-10.0, 10.0, n))
        pos = SCENE_CENTER + np.stack([r * np.cos(el) * np.cos(az), r * np.cos(el) * np.sin(az), r * np.sin(el)], 1)
        T = np.tile(np.identity(4), (n, 1, 1))
        for i in range(n):
            z = SCENE_CENTER - pos[i]; z /= np.linalg.norm(z)
            x = np.cross(z, [0.0, 0.0, 1.0]); x /= np.linalg.norm(x)
            y = np.cross(z, x)
            c, s = math.cos(roll[i]), math.sin(roll[i])
            T[i, :3, 0], T[i, :3, 1], T[i, :3, 2], T[i, :3, 3] = c * x + s * y, -s * x + c * y, z, pos[i]
        return T

    def render(self, robot2cam, seed=0, device='cpu', noise_mm=0.5, dropout=0.02, H=480, W=640, only_object=None):
        """Analytic ray casting of a batch of frames with torch (CPU or CUDA): robot2cam [F,4,4] ->
        label [F,H,W] uint8 (0 = background, k+1 = object k; `only_object` k renders object k alone as 255),
        depth [F,H,W] int16 storage of the uint16 millimetres (background 900..1100 mm, `dropout` zeros)."""
        import torch
        dev = torch.device(device)
        T = torch.as_tensor(np.asarray(robot2cam), dtype=torch.float64, device=dev)
        F = T.shape[0]
        g = torch.Generator(device=dev).manual_seed(int(seed))
        vv, uu = torch.meshgrid(torch.arange(H, device=dev, dtype=torch.float64), torch.arange(W, device=dev, dtype=torch.float64), indexing='ij')
        d_cam = torch.stack([(uu - INTR['ppx']) / INTR['fx'], (vv - INTR['ppy']) / INTR['fy'], torch.ones_like(uu)], -1)   # [H,W,3]
        best = torch.full((F, H, W), float('inf'), dtype=torch.float64, device=dev)
        label = torch.zeros((F, H, W), dtype=torch.uint8, device=dev)
        objs = range(self.n) if only_object is None else [only_object]
        for k in objs:
            A = torch.as_tensor((self.R[k] / self.axes[k]).T, dtype=torch.float64, device=dev)        # rows (R[:,i] / a_i)^T
            M = A @ T[:, :3, :3]                                                                       # [F,3,3]: robot dir -> scaled object frame
            dd = torch.einsum('fij,hwj->fhwi', M, d_cam)
            cc = torch.einsum('ij,fj->fi', A, torch.as_tensor(self.centers[k], dtype=torch.float64, device=dev) - T[:, :3, 3])
            qa = (dd * dd).sum(-1); qb = -2.0 * torch.einsum('fhwi,fi->fhw', dd, cc); qc = (cc * cc).sum(-1)[:, None, None] - 1.0
            disc = qb * qb - 4 * qa * qc
            s = torch.where(disc > 0, (-qb - torch.sqrt(disc.clamp_min(0))) / (2 * qa), torch.full_like(qa, float('inf')))
            s = torch.where(s > 0, s, torch.full_like(s, float('inf')))
            closer = s < best
            best = torch.where(closer, s, best)
            label = torch.where(closer, torch.full_like(label, 255 if only_object is not None else k + 1), label)
        hit = torch.isfinite(best)
        z = best + torch.randn((F, H, W), dtype=torch.float64, device=dev, generator=g) * noise_mm
        bg = torch.rand((F, H, W), dtype=torch.float64, device=dev, generator=g) * 200.0 + 900.0
        depth = torch.where(hit, z.clamp(1, 65535), bg).round().to(torch.int32)
        depth[torch.rand((F, H, W), device=dev, generator=g) < dropout] = 0
        depth16 = torch.where(depth >= 32768, depth - 65536, depth).to(torch.int16)
        return label, depth16
