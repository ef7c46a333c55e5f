"""Drop-in `PoseNet` / `PoseRefineNet` (reference: DenseFusion/lib/network.py:70-132, :170-206).

Same constructor arguments, same parameter names and shapes (reference checkpoints load with
`load_state_dict`, including the `cnn.model.module.*` keys of the DataParallel-wrapped encoder), same
forward signatures and return shapes.  Everything after the colour encoder runs on the hand-written
sm_100a kernels (ops.NetHandle); the PSPNet/ResNet-18 encoder stays on PyTorch/cuDNN (outside the graft,
BASELINE.json).  Extension over the reference: a batch of B objects is processed at once (the
reference silently returns element 0 only, network.py:123).

Training (row a16, DenseFusion/tools/train.py:215-233): when autograd is recording, `PoseRefineNet.forward` runs the
bf16 training forward of csrc/train.cuh and its backward accumulates into the parameters' `.grad` through the
hand-written dgrad / wgrad kernels -- `dis.backward()` and `torch.optim.Adam(refiner.parameters())` work as in the
reference.  The parameters are views of ONE flat fp32 vector (and the gradients of one flat gradient vector, which
is what gets all-reduced over NCCL between ranks).  Training the ESTIMATOR (`PoseNet` with autograd) is outside the
graft (SURVEY 8 row a16 covers the refiner step; the estimator is in eval mode there, train.py:191-193) and raises
unless `PoseNet.allow_torch_training` is set, in which case the layer stack runs as torch ops on the GPU like the
colour encoder does.  CPU tensors are always an error: there is no CPU path.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops


# ----------------------------------------------------------------------------------------------------
# Colour encoder (outside the graft): ResNet-18 (dilated, no BatchNorm) + pyramid pooling + 3 upsampling
# stages -> 32 channels per pixel.  Key names follow DenseFusion/lib/pspnet.py and lib/extractors.py.
class _Block(nn.Module):
    def __init__(self, cin, cout, stride=1, dilation=1, project=False):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride=stride, padding=dilation, dilation=dilation, bias=False)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=dilation, dilation=dilation, bias=False)
        self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride, bias=False)) if project else None

    def forward(self, x):
        y = self.conv2(F.relu(self.conv1(x)))
        return F.relu(y + (x if self.downsample is None else self.downsample(x)))


class _ResNet18(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
        # (cin, cout, stride of the first block, dilation of the following block)
        spec = [(64, 64, 1, 1), (64, 128, 2, 1), (128, 256, 1, 2), (256, 512, 1, 4)]
        for i, (ci, co, st, dil) in enumerate(spec, 1):
            first = _Block(ci, co, stride=st, dilation=1, project=(st != 1 or ci != co))
            setattr(self, 'layer%d' % i, nn.Sequential(first, _Block(co, co, dilation=dil)))

    def forward(self, x):
        x = self.maxpool(F.relu(self.conv1(x)))
        return self.layer4(self.layer3(self.layer2(self.layer1(x))))


class _Pyramid(nn.Module):
    def __init__(self, feats=512, out=1024, sizes=(1, 2, 3, 6)):
        super().__init__()
        self.stages = nn.ModuleList(nn.Sequential(nn.AdaptiveAvgPool2d((s, s)), nn.Conv2d(feats, feats, 1, bias=False)) for s in sizes)
        self.bottleneck = nn.Conv2d(feats * (len(sizes) + 1), out, 1)

    def forward(self, f):
        hw = f.shape[2:]
        pri = [F.interpolate(st(f), size=hw, mode='bilinear', align_corners=False) for st in self.stages] + [f]
        return F.relu(self.bottleneck(torch.cat(pri, 1)))


class _Up(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Sequential(nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True),
                                  nn.Conv2d(cin, cout, 3, padding=1), nn.PReLU())

    def forward(self, x):
        return self.conv(x)


class _PSPNet(nn.Module):
    def __init__(self, n_classes=21):
        super().__init__()
        self.feats = _ResNet18()
        self.psp = _Pyramid(512, 1024)
        self.drop_1 = nn.Dropout2d(p=0.3)
        self.up_1, self.up_2, self.up_3 = _Up(1024, 256), _Up(256, 64), _Up(64, 64)
        self.drop_2 = nn.Dropout2d(p=0.15)
        self.final = nn.Sequential(nn.Conv2d(64, 32, 1), nn.LogSoftmax(dim=1))
        self.classifier = nn.Sequential(nn.Linear(256, 256), nn.ReLU(), nn.Linear(256, n_classes))   # unused by forward (as upstream)

    def forward(self, x):
        p = self.drop_1(self.psp(self.feats(x)))
        p = self.drop_2(self.up_1(p))
        p = self.drop_2(self.up_2(p))
        return self.final(self.up_3(p))


class _Holder(nn.Module):
    """Gives the encoder the `model.module.` key prefix of the reference's nn.DataParallel wrapper."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, x):
        return self.module(x)


class ModifiedResnet(nn.Module):
    def __init__(self, usegpu=True):
        super().__init__()
        self.model = _Holder(_PSPNet())

    def forward(self, x):
        return self.model(x)


# ----------------------------------------------------------------------------------------------------
class PoseNetFeat(nn.Module):
    """Parameter holder with the reference's names (network.py:39-51); computed by the fused trunk."""

    def __init__(self, num_points, refine=False):
        super().__init__()
        self.conv1 = nn.Conv1d(3, 64, 1); self.conv2 = nn.Conv1d(64, 128, 1)
        self.e_conv1 = nn.Conv1d(32, 64, 1); self.e_conv2 = nn.Conv1d(64, 128, 1)
        self.conv5 = nn.Conv1d(384 if refine else 256, 512, 1); self.conv6 = nn.Conv1d(512, 1024, 1)
        self.num_points, self.refine = num_points, refine

    def forward(self, x, emb):                     # torch path (autograd / training only)
        x1 = F.relu(self.conv1(x)); e1 = F.relu(self.e_conv1(emb))
        x2 = F.relu(self.conv2(x1)); e2 = F.relu(self.e_conv2(e1))
        pf1, pf2 = torch.cat((x1, e1), 1), torch.cat((x2, e2), 1)
        y = F.relu(self.conv6(F.relu(self.conv5(torch.cat((pf1, pf2), 1) if self.refine else pf2))))
        ap = y.mean(dim=2, keepdim=True)
        return ap.view(-1, 1024) if self.refine else torch.cat([pf1, pf2, ap.expand(-1, -1, x.shape[2])], 1)


PoseRefineNetFeat = lambda num_points: PoseNetFeat(num_points, refine=True)   # noqa: E731  (network.py:136)


class _Grafted(nn.Module):
    """Builds / refreshes the C-side handle (split-bf16 weights + workspace) from the current parameters."""
    _kind = None

    def _handle(self, batch, n_points):
        ps = [p for n, p in self.named_parameters() if not n.startswith('cnn.')]
        key = (tuple(p._version for p in ps), tuple(p.data_ptr() for p in ps), str(ps[0].device))
        h = getattr(self, '_ape', None)
        if h is None or self._ape_key != key or h.max_batch < batch or h.max_points < n_points:
            if h is not None:
                h.close()
            sd = {k: v for k, v in self.state_dict().items() if not k.startswith('cnn.')}
            self._ape = ops.NetHandle(self._kind, sd, self.num_obj, max(batch, getattr(h, 'max_batch', 1)),
                                      max(n_points, getattr(h, 'max_points', 1)), device=ps[0].device)
            self._ape_key = key
        return self._ape

    def _needs_autograd(self):
        return torch.is_grad_enabled() and any(p.requires_grad for n, p in self.named_parameters() if not n.startswith('cnn.')) \
            and self.training


class PoseNet(_Grafted):
    _kind = ops.NET_POSENET

    def __init__(self, num_points, num_obj):
        super().__init__()
        self.num_points, self.num_obj = num_points, num_obj
        self.cnn = ModifiedResnet()
        self.feat = PoseNetFeat(num_points)
        for h, w in (('r', 4), ('t', 3), ('c', 1)):
            setattr(self, 'conv1_' + h, nn.Conv1d(1408, 640, 1)); setattr(self, 'conv2_' + h, nn.Conv1d(640, 256, 1))
            setattr(self, 'conv3_' + h, nn.Conv1d(256, 128, 1)); setattr(self, 'conv4_' + h, nn.Conv1d(128, num_obj * w, 1))

    def forward(self, img, x, choose, obj):
        """img [B,3,h,w], x [B,N,3], choose [B,1,N] int64 (flat indices into h*w), obj [B,1] int64
        -> pred_r [B,N,4], pred_t [B,N,3], pred_c [B,N,1], emb [B,32,N] (detached), as network.py:95-132."""
        out_img = self.cnn(img)                                        # PyTorch/cuDNN, outside the graft
        return self.forward_geometry(out_img, x, choose, obj)

    allow_torch_training = False      # opt-in: estimator training as torch ops on the GPU (outside the graft)

    def forward_geometry(self, out_img, x, choose, obj):
        """Everything after the encoder (network.py:98-132) on the sm_100a kernels."""
        B, N = x.shape[0], x.shape[1]
        if not out_img.is_cuda:
            raise ops._lib.ApeError('PoseNet: tensors must be on a CUDA device (no CPU fallback)')
        if self._needs_autograd():
            if not self.allow_torch_training:
                raise NotImplementedError('PoseNet training is outside the graft (SURVEY 8: only the refiner training step, '
                                          'train.py:215-233, is grafted); call .eval() / torch.no_grad() for inference or set '
                                          'PoseNet.allow_torch_training = True to run the estimator as torch ops on the GPU')
            return self._torch_geometry(out_img, x, choose, obj)
        return self._handle(B, N).posenet_forward(out_img.detach(), x, choose, obj)

    def _torch_geometry(self, out_img, x, choose, obj):               # estimator training only; outside the graft
        B, di = out_img.shape[:2]
        N = x.shape[1]
        emb = torch.gather(out_img.reshape(B, di, -1), 2, choose.reshape(B, 1, N).expand(B, di, N)).contiguous()
        ap = self.feat(x.transpose(2, 1).contiguous(), emb)
        outs = []
        for h, w in (('r', 4), ('t', 3), ('c', 1)):
            y = ap
            for i in (1, 2, 3):
                y = F.relu(getattr(self, 'conv%d_%s' % (i, h))(y))
            y = getattr(self, 'conv4_' + h)(y)
            y = (torch.sigmoid(y) if h == 'c' else y).view(B, self.num_obj, w, N)
            sel = y[torch.arange(B, device=y.device), obj.reshape(B)]               # [B,w,N]
            outs.append(sel.transpose(2, 1).contiguous())
        return outs[0], outs[1], outs[2], emb.detach()


class _RefinerTrainFn(torch.autograd.Function):
    """Autograd node of the training forward: backward accumulates straight into the flat gradient vector that the
    parameters' `.grad` tensors are views of (so nothing is returned for the parameter inputs)."""

    @staticmethod
    def forward(ctx, module, x, emb, obj, *params):
        tr = module._trainer(x.shape[0], x.shape[1])
        r, t = tr.forward(x, emb, obj)
        module._fwd_id += 1                       # the trainer workspace now holds THIS forward's activations
        ctx.module, ctx.fwd_id, ctx.tr = module, module._fwd_id, tr
        ctx.save_for_backward(x, emb, obj)
        return r, t

    @staticmethod
    def backward(ctx, d_r, d_t):
        x, emb, obj = ctx.saved_tensors
        m = ctx.module
        tr = m._trainer(x.shape[0], x.shape[1])
        if m._fwd_id != ctx.fwd_id or tr is not ctx.tr:
            # another forward ran in between (e.g. `(dis_0 + dis_1).backward()`, a loss summed over several objects):
            # the workspace holds ITS activations, so the forward being differentiated is re-run from the saved inputs
            # first (same kernels, same result) -- the patterns autograd supports stay correct, at the cost of a recompute
            tr.forward(x, emb, obj)
            m._fwd_id += 1
        m._attach_grads()
        tr.backward(x, emb, obj, d_r.contiguous(), d_t.contiguous())
        return (None,) * (4 + len(list(m.parameters())))


class PoseRefineNet(_Grafted):
    _kind = ops.NET_REFINER

    def __init__(self, num_points, num_obj):
        super().__init__()
        self.num_points, self.num_obj = num_points, num_obj
        self.feat = PoseNetFeat(num_points, refine=True)
        self.conv1_r, self.conv1_t = nn.Linear(1024, 512), nn.Linear(1024, 512)
        self.conv2_r, self.conv2_t = nn.Linear(512, 128), nn.Linear(512, 128)
        self.conv3_r, self.conv3_t = nn.Linear(128, num_obj * 4), nn.Linear(128, num_obj * 3)
        self._tr = None
        self._fwd_id = 0

    # -- training plumbing -------------------------------------------------------------------------
    def _trainer(self, batch, n_points):
        """Trainer handle whose flat parameter vector the module's parameters alias (re-created when the workspace
        must grow); bf16 weight copies are re-derived whenever a parameter was updated in place (optimizer step)."""
        tr = self._tr
        dev = next(self.parameters()).device
        first_name, first = next(iter(self.named_parameters()))
        moved = tr is not None and first.data_ptr() != tr.view(first_name).data_ptr()      # .to() / .cuda() re-allocated them
        if tr is None or moved or tr.max_batch < batch or tr.max_points < n_points or tr.device != dev:
            grads = None if (tr is None or moved) else tr.grads.clone()
            sd = {k: v.detach() for k, v in self.state_dict().items()}
            with torch.cuda.device(dev):
                new = ops.RefinerTrainerHandle(sd, self.num_obj, max(batch, getattr(tr, 'max_batch', 1)),
                                               max(n_points, getattr(tr, 'max_points', 1)), device=dev)
            if grads is not None:
                new.grads.copy_(grads)
                tr.close()
            for name, p in self.named_parameters():
                had_grad = p.grad is not None
                p.data = new.view(name)
                p.grad = new.view(name, new.grads) if had_grad else None
            self._tr = tr = new
            self._tr_key = None
        key = tuple(p._version for p in self.parameters())
        if key != self._tr_key:
            tr.sync_weights()
            self._tr_key = key
        return tr

    def _attach_grads(self):
        """Make every parameter's .grad a view of the flat gradient vector (zeroed first when the optimizer dropped
        the gradients with zero_grad(set_to_none=True))."""
        tr = self._tr
        params = list(self.named_parameters())
        if all(p.grad is None for _, p in params):
            tr.grads.zero_()
        for name, p in params:
            view = tr.view(name, tr.grads)
            if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                if p.grad is not None:
                    view.copy_(p.grad)
                p.grad = view

    def flat_gradient(self):
        """The flat fp32 gradient vector (all-reduce this between ranks), or None before the first backward."""
        return None if self._tr is None else self._tr.grads

    def forward(self, x, emb, obj):
        """x = new_points [B,N,3], emb [B,32,N], obj [B,1] -> pred_r [B,4], pred_t [B,3] (network.py:187-206)."""
        B, N = x.shape[0], x.shape[1]
        if not x.is_cuda:
            raise ops._lib.ApeError('PoseRefineNet: tensors must be on a CUDA device (no CPU fallback)')
        if self._needs_autograd():
            return _RefinerTrainFn.apply(self, x.detach(), emb.detach(), obj, *self.parameters())
        # inference always runs the split-bf16 kernels on their own handle / workspace (re-derived from the parameters when
        # an optimizer step changed them): it neither disturbs the activations a pending backward needs nor drops to the
        # plain-bf16 accuracy of the training forward
        return self._handle(B, N).refiner_forward(x, emb, obj)
