"""Hard NMS and stable top-k.

Mirrors mxdetection/ops NMS (/root/reference/README.md:24) with the semantics
of ``mx.nd.contrib.box_nms`` (mxnet 1.3.0): stable score-descending order
(ties -> lower index), strict ``iou > thr`` (SURVEY.md 8(a) Spec B).
"""
import math

import torch

from .. import _lib as L

_FMT = {"corner": 0, "center": 1}


def topk_stable(scores, k):
    """scores (n) or (S,n) -> (idx int32, vals f32) of the top-k by (score desc, index asc)."""
    L.require_cuda(scores)
    scores = scores.contiguous()
    n = scores.shape[-1]
    kk = k if 0 < k < n else n
    shape = tuple(scores.shape[:-1]) + (kk,)
    idx = torch.empty(shape, dtype=torch.int32, device=scores.device)
    vals = torch.empty(shape, dtype=torch.float32, device=scores.device)
    segs = int(scores.numel() // max(n, 1))
    nbytes = L.lib.mxd_topk_stable_workspace_bytes(segs, int(n), int(k))     # > 0 only above MXD_SORT_CAP rows
    ws = L.workspace(nbytes, scores.device, "topk") if nbytes else None
    L.call("mxd_topk_stable", L.dl(scores), L.dl(idx), L.dl(vals), int(k), ws.data_ptr() if nbytes else None,
           ws.numel() if nbytes else 0, L.current_stream(scores.device))
    return idx, vals


def nms_indices(boxes, scores, iou_thr, delta=0.0, topk=-1, valid_thresh=-math.inf, ids=None, force_suppress=True,
                max_out=-1):
    """Returns (keep int32 (cap) padded with -1, num_keep int32 (1)) - no host sync."""
    L.require_cuda(boxes, scores, ids)
    boxes = boxes.contiguous(); scores = scores.contiguous()
    n = boxes.shape[0]
    k = topk if 0 < topk < n else n
    cap = min(k, max_out) if max_out > 0 else k
    keep = torch.empty((max(cap, 0),), dtype=torch.int32, device=boxes.device)      # the library pads with -1
    num = torch.empty((1,), dtype=torch.int32, device=boxes.device)
    nbytes = L.lib.mxd_nms_workspace_bytes(int(n), int(topk))
    ws = L.workspace(nbytes, boxes.device, "nms")
    L.call("mxd_nms", L.dl(boxes), L.dl(scores), L.dl(ids), L.dl(keep), L.dl(num), float(iou_thr), float(delta),
           int(topk), float(valid_thresh), 1 if force_suppress else 0, int(max_out), ws.data_ptr(), ws.numel(),
           L.current_stream(boxes.device))
    return keep, num


def nms_batched(boxes, scores, seg_offsets, max_seg_len, iou_thr, delta=0.0, topk=-1, valid_thresh=-math.inf, ids=None,
                force_suppress=True, max_out=-1):
    """NMS over ragged segments of one box array (segment s = rows seg_offsets[s] .. seg_offsets[s+1], int32 on the
    device; max_seg_len a host upper bound).  Returns (keep (S,cap) int32 GLOBAL row indices, -1 padded; num_keep (S))."""
    L.require_cuda(boxes, scores, seg_offsets, ids)
    boxes = boxes.contiguous(); scores = scores.contiguous(); seg_offsets = seg_offsets.contiguous()
    S = seg_offsets.shape[0] - 1
    k = topk if 0 < topk < max_seg_len else max_seg_len
    cap = min(k, max_out) if max_out > 0 else k
    keep = torch.empty((S, max(cap, 0)), dtype=torch.int32, device=boxes.device)
    num = torch.empty((S,), dtype=torch.int32, device=boxes.device)
    nbytes = L.lib.mxd_nms_batched_workspace_bytes(int(S), int(max_seg_len), int(topk))
    ws = L.workspace(nbytes, boxes.device, "nms_batched")
    L.call("mxd_nms_batched", L.dl(boxes), L.dl(scores), L.dl(ids), L.dl(seg_offsets), int(max_seg_len), L.dl(keep),
           L.dl(num), float(iou_thr), float(delta), int(topk), float(valid_thresh), 1 if force_suppress else 0,
           int(max_out), ws.data_ptr(), ws.numel(), L.current_stream(boxes.device))
    return keep, num


def nms(dets, iou_thr, delta=1.0, max_out=-1):
    """mmdet-0.5 style: dets (n,5) [x1,y1,x2,y2,score] -> (dets[keep], keep int64).  Host-syncs for the count."""
    keep, num = nms_indices(dets[:, :4].contiguous(), dets[:, 4].contiguous(), iou_thr, delta=delta, max_out=max_out)
    keep = keep[: int(num.item())].long()
    return dets[keep], keep


def box_nms(data, overlap_thresh=0.5, valid_thresh=0.0, topk=-1, coord_start=2, score_index=1, id_index=-1,
            force_suppress=False, in_format="corner", out_format="corner", return_index=False):
    """Drop-in for mx.nd.contrib.box_nms: (..., N, K) -> same shape, kept rows first, others -1."""
    L.require_cuda(data)
    data = data.contiguous()
    shape = data.shape
    x = data.reshape(-1, shape[-2], shape[-1])
    out = torch.empty_like(x)
    index = torch.empty(x.shape[:2], dtype=torch.int32, device=x.device)
    nbytes = L.lib.mxd_box_nms_workspace_bytes(int(x.shape[0]), int(x.shape[1]), int(topk))
    ws = L.workspace(nbytes, x.device, "box_nms")
    L.call("mxd_box_nms", L.dl(x), L.dl(out), L.dl(index), float(overlap_thresh), float(valid_thresh), int(topk),
           int(coord_start), int(score_index), int(id_index), 1 if force_suppress else 0, _FMT[in_format],
           _FMT[out_format], ws.data_ptr(), ws.numel(), L.current_stream(x.device))
    out = out.reshape(shape)
    if return_index:
        return out, index.reshape(shape[:-1])
    return out


def box_nms_backward(out_grad, index):
    """_backward_box_nms: routes out_grad rows back to their source rows."""
    L.require_cuda(out_grad, index)
    shape = out_grad.shape
    g = out_grad.contiguous().reshape(-1, shape[-2], shape[-1])
    ig = torch.empty_like(g)
    L.call("mxd_box_nms_backward", L.dl(g), L.dl(index.contiguous().reshape(g.shape[0], g.shape[1])), L.dl(ig),
           L.current_stream(g.device))
    return ig.reshape(shape)
