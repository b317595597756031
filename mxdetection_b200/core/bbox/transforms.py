"""Delta encode / decode (SURVEY.md 8(a) Spec F; mmdet-0.5 bbox2delta / delta2bbox)."""
import torch

from ... import _lib as L


def bbox2delta(proposals, gt, means=(0, 0, 0, 0), stds=(1, 1, 1, 1)):
    L.require_cuda(proposals, gt)
    out = torch.empty_like(proposals, dtype=torch.float32)
    L.call("mxd_bbox2delta", L.dl(proposals.contiguous()), L.dl(gt.contiguous()), L.dl(out), L.float4(means),
           L.float4(stds), L.current_stream(out.device))
    return out


def delta2bbox(rois, deltas, means=(0, 0, 0, 0), stds=(1, 1, 1, 1), max_shape=None, wh_ratio_clip=16 / 1000):
    """max_shape = (h, w) clips x to [0,w-1], y to [0,h-1]."""
    L.require_cuda(rois, deltas)
    out = torch.empty_like(rois, dtype=torch.float32)
    mh, mw = (int(max_shape[0]), int(max_shape[1])) if max_shape is not None else (0, 0)
    L.call("mxd_delta2bbox", L.dl(rois.contiguous()), L.dl(deltas.contiguous()), L.dl(out), L.float4(means),
           L.float4(stds), mh, mw, float(wh_ratio_clip), L.current_stream(out.device))
    return out


def bbox2roi(bbox_list):
    """List of per-image (n,4+) boxes -> (sum n, 5) [batch_ind, x1, y1, x2, y2]."""
    rois = []
    for i, b in enumerate(bbox_list):
        ind = torch.full((b.shape[0], 1), float(i), dtype=torch.float32, device=b.device)
        rois.append(torch.cat([ind, b[:, :4].float()], dim=1))
    return torch.cat(rois, 0)
