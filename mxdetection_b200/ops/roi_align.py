"""RoIAlign forward / backward.

Mirrors ``mx.nd.contrib.ROIAlign(data, rois, pooled_size, spatial_scale,
sample_ratio=-1)`` of mxnet 1.3.0 (/root/reference/README.md:37) as exposed by
mxdetection/ops (/root/reference/README.md:24), and the mmdet-0.5 style
``RoIAlign(out_size, spatial_scale, sample_num)`` layer built on it.
Semantics: SURVEY.md 8(a) Spec A.
"""
import ctypes

import torch

from .. import _lib as L


def _pair(v):
    return (int(v), int(v)) if isinstance(v, int) else (int(v[0]), int(v[1]))


USE_PLANE_KERNELS = True   # False: pass workspace=NULL so the library uses its gather kernels (A/B testing)


def _ws(num_rois, shapes, ph, pw, sample_ratio, device):
    """Scratch for the plane-resident kernels (RoI plan + packed tap tables)."""
    if not USE_PLANE_KERNELS:
        return None, 0
    n_lv = len(shapes)
    hs = (ctypes.c_int * n_lv)(*[int(s[2]) for s in shapes])
    ws_ = (ctypes.c_int * n_lv)(*[int(s[3]) for s in shapes])
    nbytes = L.lib.mxd_roi_align_workspace_bytes(int(num_rois), int(shapes[0][0]), int(shapes[0][1]), n_lv, hs, ws_,
                                                 ph, pw, int(sample_ratio))
    buf = L.workspace(nbytes, device, "roi_align")
    return buf.data_ptr(), buf.numel()


def roi_align_forward(data, rois, pooled_size, spatial_scale, sample_ratio=-1, out=None):
    """data (N,C,H,W) f32 cuda, rois (R,5) [batch,x1,y1,x2,y2] -> (R,C,PH,PW)."""
    L.require_cuda(data, rois)
    ph, pw = _pair(pooled_size)
    data = data.contiguous(); rois = rois.contiguous()
    if out is None:
        out = torch.empty((rois.shape[0], data.shape[1], ph, pw), dtype=torch.float32, device=data.device)
    wp, wn = _ws(rois.shape[0], [data.shape], ph, pw, sample_ratio, data.device)
    L.call("mxd_roi_align_forward", L.dl(data), L.dl(rois), L.dl(out), ph, pw, float(spatial_scale),
           int(sample_ratio), wp, wn, L.current_stream(data.device))
    return out


def roi_align_backward(grad_out, rois, data_shape, pooled_size, spatial_scale, sample_ratio=-1, grad_data=None,
                       accumulate=False):
    """Returns grad wrt data.  grad_data: optional destination; accumulate=True is req='add'
    (added in place), False is req='write' (the library zero-fills it first)."""
    L.require_cuda(grad_out, rois)
    ph, pw = _pair(pooled_size)
    grad_out = grad_out.contiguous(); rois = rois.contiguous()
    if grad_data is None:
        accumulate = False
        grad_data = torch.empty(tuple(data_shape), dtype=torch.float32, device=grad_out.device)
    wp, wn = _ws(rois.shape[0], [grad_data.shape], ph, pw, sample_ratio, grad_out.device)
    L.call("mxd_roi_align_backward", L.dl(grad_out), L.dl(rois), L.dl(grad_data), ph, pw, float(spatial_scale),
           int(sample_ratio), 1 if accumulate else 0, wp, wn, L.current_stream(grad_out.device))
    return grad_data


def _scales(spatial_scales):
    return (ctypes.c_float * len(spatial_scales))(*[float(s) for s in spatial_scales])


def roi_align_fpn_forward(feats, rois, pooled_size, spatial_scales, sample_ratio=2, levels=None, finest_scale=56,
                          out=None):
    """Multi-level RoIAlign in one launch: feats[l] (N,C,H_l,W_l); levels (R) i32 or None (Spec G in-kernel)."""
    L.require_cuda(rois, *feats)
    ph, pw = _pair(pooled_size)
    feats = [f.contiguous() for f in feats]
    rois = rois.contiguous()
    if out is None:
        out = torch.empty((rois.shape[0], feats[0].shape[1], ph, pw), dtype=torch.float32, device=rois.device)
    arr, keep = L.dl_array(feats)
    wp, wn = _ws(rois.shape[0], [f.shape for f in feats], ph, pw, sample_ratio, rois.device)
    L.call("mxd_roi_align_fpn_forward", arr, len(feats), _scales(spatial_scales), L.dl(rois), L.dl(levels),
           L.dl(out), ph, pw, int(sample_ratio), float(finest_scale), wp, wn, L.current_stream(rois.device))
    del keep
    return out


def roi_align_fpn_backward(grad_out, rois, feat_shapes, pooled_size, spatial_scales, sample_ratio=2, levels=None,
                           finest_scale=56, grad_feats=None, accumulate=False):
    """grad_feats: optional destinations; accumulate=True adds into them (req='add')."""
    L.require_cuda(grad_out, rois)
    ph, pw = _pair(pooled_size)
    grad_out = grad_out.contiguous(); rois = rois.contiguous()
    if grad_feats is None:
        accumulate = False
        grad_feats = [torch.empty(tuple(s), dtype=torch.float32, device=grad_out.device) for s in feat_shapes]
    arr, keep = L.dl_array(grad_feats)
    wp, wn = _ws(rois.shape[0], [g.shape for g in grad_feats], ph, pw, sample_ratio, rois.device)
    L.call("mxd_roi_align_fpn_backward", L.dl(grad_out), L.dl(rois), L.dl(levels), arr, len(grad_feats),
           _scales(spatial_scales), ph, pw, int(sample_ratio), float(finest_scale), 1 if accumulate else 0,
           wp, wn, L.current_stream(rois.device))
    del keep
    return grad_feats


class RoIAlignFunction(torch.autograd.Function):
    """Autograd glue (MXNet's engine is not available; torch supplies the tape)."""

    @staticmethod
    def forward(ctx, data, rois, pooled_size, spatial_scale, sample_ratio):
        ctx.save_for_backward(rois)
        ctx.cfg = (tuple(data.shape), _pair(pooled_size), float(spatial_scale), int(sample_ratio))
        return roi_align_forward(data, rois, pooled_size, spatial_scale, sample_ratio)

    @staticmethod
    def backward(ctx, grad_out):
        (rois,) = ctx.saved_tensors
        shape, ps, scale, sr = ctx.cfg
        gd = roi_align_backward(grad_out, rois, shape, ps, scale, sr)
        return gd, torch.zeros_like(rois), None, None, None   # grad_rois == 0 (Spec A)


def ROIAlign(data, rois, pooled_size, spatial_scale, sample_ratio=-1):
    """Drop-in for mx.nd.contrib.ROIAlign (differentiable wrt data)."""
    return RoIAlignFunction.apply(data, rois, pooled_size, spatial_scale, sample_ratio)


class RoIAlign(torch.nn.Module):
    """mmdet-0.5 style layer: RoIAlign(out_size, spatial_scale, sample_num)."""

    def __init__(self, out_size, spatial_scale, sample_num=0):
        super().__init__()
        self.out_size = _pair(out_size)
        self.spatial_scale = float(spatial_scale)
        self.sample_num = int(sample_num)

    def forward(self, features, rois):
        sr = self.sample_num if self.sample_num > 0 else -1
        return RoIAlignFunction.apply(features, rois, self.out_size, self.spatial_scale, sr)
