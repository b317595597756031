"""FPN RoI feature extractor (SURVEY.md 8(a) Spec G + row G2; mmdet-0.5
SingleLevelRoI of mxdetection/models/roi_extractors, /root/reference/README.md:32).

The reference style is one RoIAlign launch per level over boolean-masked RoI
subsets plus a scatter; here every RoI carries its level and ONE launch reads
from a pointer table of the level maps."""
import torch

from ... import _lib as L
from ...ops.roi_align import roi_align_fpn_backward, roi_align_fpn_forward


def map_roi_levels(rois, num_levels, finest_scale=56):
    """rois (R,5) or (R,4) -> level index (R) int32: scale <112 -> 0, <224 -> 1, <448 -> 2, else 3 (for 4 levels)."""
    L.require_cuda(rois)
    out = torch.empty((rois.shape[0],), dtype=torch.int32, device=rois.device)
    L.call("mxd_map_roi_levels", L.dl(rois.contiguous()), L.dl(out), int(num_levels), float(finest_scale),
           L.current_stream(rois.device))
    return out


class _FpnRoIAlignFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rois, levels, cfg, *feats):
        out_size, scales, sample_num, finest = cfg
        ctx.cfg = cfg
        ctx.shapes = [tuple(f.shape) for f in feats]
        ctx.save_for_backward(rois, levels)
        return roi_align_fpn_forward(feats, rois, out_size, scales, sample_num, levels, finest)

    @staticmethod
    def backward(ctx, grad_out):
        rois, levels = ctx.saved_tensors
        out_size, scales, sample_num, finest = ctx.cfg
        grads = roi_align_fpn_backward(grad_out, rois, ctx.shapes, out_size, scales, sample_num, levels, finest)
        return (None, None, None) + tuple(grads)


class SingleLevelRoI(torch.nn.Module):
    """SingleLevelRoI(out_size, featmap_strides, sample_num=2, finest_scale=56)."""

    def __init__(self, out_size=7, featmap_strides=(4, 8, 16, 32), sample_num=2, finest_scale=56):
        super().__init__()
        self.out_size = (out_size, out_size) if isinstance(out_size, int) else tuple(out_size)
        self.featmap_strides = tuple(featmap_strides)
        self.sample_num = int(sample_num)
        self.finest_scale = finest_scale

    @property
    def num_inputs(self):
        return len(self.featmap_strides)

    def map_roi_levels(self, rois, num_levels):
        return map_roi_levels(rois, num_levels, self.finest_scale)

    def forward(self, feats, rois):
        feats = list(feats[: self.num_inputs])
        levels = self.map_roi_levels(rois, len(feats))
        scales = tuple(1.0 / s for s in self.featmap_strides[: len(feats)])
        cfg = (self.out_size, scales, self.sample_num if self.sample_num > 0 else -1, float(self.finest_scale))
        return _FpnRoIAlignFunction.apply(rois, levels, cfg, *feats)
