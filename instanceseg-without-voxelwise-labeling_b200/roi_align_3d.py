"""RoIAlign3D behind the reference names
(lib/modeling/roi_xfrom/roi_align_3d/functions/roi_align_3d.py:7-51, modules/roi_align_3d.py:6-48).

`RoIAlignFunction_3d(P_s, P_h, P_w, scale, sr)(features, rois)` keeps the reference's
constructor-then-call form (model_builder.py:280-281, 311-312) on top of a modern
torch.autograd.Function.  layout="reference" (default) reproduces the reference bit layout:
the forward emits bins in (H,W,S) order into the [R,C,Ps,Ph,Pw] tensor while the backward reads
grad in (S,H,W) order with the -0.1 z guard; layout="shw" is the self-consistent variant whose
backward is the exact adjoint of its forward."""
import torch
from torch.autograd import Function
from torch.nn.functional import avg_pool3d, max_pool3d
from torch.nn.modules.module import Module

from . import _lib

_DT = {torch.float32: 0, torch.bfloat16: 1}


def roialign3d_forward(features, rois, Ps, Ph, Pw, scale, sr, layout=0):
    if not features.is_cuda:
        raise NotImplementedError            # functions/roi_align_3d.py:31-32
    if features.dtype not in _DT:
        raise TypeError("RoIAlign3D supports float32 and bfloat16 features, got %s" % features.dtype)
    if rois.dim() != 2 or rois.size(1) != 7:
        raise ValueError("rois must be [R,7] (batch,x1,y1,z1,x2,y2,z2)")
    features = features.contiguous()
    rois = rois.to(device=features.device, dtype=torch.float32).contiguous()
    B, Cc, S, H, W = features.shape
    R = rois.size(0)
    out = torch.empty((R, Cc, Ps, Ph, Pw), dtype=features.dtype, device=features.device)
    if R:
        _lib.check(_lib.lib().b200seg_roialign3d_fwd_dev(
            _lib.ptr(features), _DT[features.dtype], _lib.ptr(rois), _lib.ptr(out), B, Cc, S, H, W, R,
            Ps, Ph, Pw, float(scale), int(sr), int(layout), _lib.current_stream()), "roialign3d_fwd")
    return out


def roialign3d_backward(grad_out, rois, feature_size, scale, sr, layout=0):
    B, Cc, S, H, W = feature_size
    grad_out = grad_out.contiguous()
    R, _, Ps, Ph, Pw = grad_out.shape
    L = _lib.lib()
    grad_in = torch.empty((B, Cc, S, H, W), dtype=grad_out.dtype, device=grad_out.device)
    ws_bytes = L.b200seg_roialign3d_bwd_workspace_bytes(R, Cc, S, H, W, max(Ps, Ph, Pw))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=grad_out.device)
    _lib.check(L.b200seg_roialign3d_bwd_dev(
        _lib.ptr(grad_out), _DT[grad_out.dtype], _lib.ptr(rois), _lib.ptr(grad_in), B, Cc, S, H, W, R,
        Ps, Ph, Pw, float(scale), int(sr), int(layout), _lib.ptr(ws), ws_bytes, _lib.current_stream()),
        "roialign3d_bwd")
    return grad_in


class _RoIAlign3D(Function):
    @staticmethod
    def forward(ctx, features, rois, Ps, Ph, Pw, scale, sr, layout):
        rois = rois.to(device=features.device, dtype=torch.float32).contiguous()
        ctx.save_for_backward(rois)
        ctx.cfg = (tuple(features.shape), Ps, Ph, Pw, scale, sr, layout)
        return roialign3d_forward(features, rois, Ps, Ph, Pw, scale, sr, layout)

    @staticmethod
    def backward(ctx, grad_output):
        (rois,) = ctx.saved_tensors
        fsize, Ps, Ph, Pw, scale, sr, layout = ctx.cfg
        assert grad_output.is_cuda
        gi = roialign3d_backward(grad_output, rois, fsize, scale, sr, layout)
        return gi, None, None, None, None, None, None, None


class RoIAlignFunction_3d(object):
    """Drop-in for the reference's old-style stateful Function: construct, then call."""

    def __init__(self, aligned_slices, aligned_height, aligned_width, spatial_scale, sampling_ratio,
                 layout="reference"):
        self.aligned_slices = int(aligned_slices)
        self.aligned_width = int(aligned_width)
        self.aligned_height = int(aligned_height)
        self.spatial_scale = float(spatial_scale)
        self.sampling_ratio = int(sampling_ratio)
        self.layout = {"reference": 0, "shw": 1}[layout]

    def __call__(self, features, rois):
        return _RoIAlign3D.apply(features, rois, self.aligned_slices, self.aligned_height, self.aligned_width,
                                 self.spatial_scale, self.sampling_ratio, self.layout)

    forward = __call__


class RoIAlign_3d(Module):
    def __init__(self, aligned_slices, aligned_height, aligned_width, spatial_scale, sampling_ratio):
        super(RoIAlign_3d, self).__init__()
        self.aligned_width = int(aligned_width)
        self.aligned_height = int(aligned_height)
        self.aligned_slices = int(aligned_slices)
        self.spatial_scale = float(spatial_scale)
        self.sampling_ratio = int(sampling_ratio)

    def forward(self, features, rois):
        return RoIAlignFunction_3d(self.aligned_slices, self.aligned_height, self.aligned_width,
                                   self.spatial_scale, self.sampling_ratio)(features, rois)


class RoIAlignAvg_3d(RoIAlign_3d):
    def forward(self, features, rois):
        x = RoIAlignFunction_3d(self.aligned_slices + 1, self.aligned_height + 1, self.aligned_width + 1,
                                self.spatial_scale, self.sampling_ratio)(features, rois)
        return avg_pool3d(x, kernel_size=2, stride=1)


class RoIAlignMax_3d(RoIAlign_3d):
    def forward(self, features, rois):
        x = RoIAlignFunction_3d(self.aligned_slices + 1, self.aligned_height + 1, self.aligned_width + 1,
                                self.spatial_scale, self.sampling_ratio)(features, rois)
        return max_pool3d(x, kernel_size=2, stride=1)
