"""Per-class score threshold -> class-aware NMS -> top max_num (mmdet-0.5 ``multiclass_nms`` /
``BBoxHead.get_det_bboxes``).  One native call (``mxd_det_bboxes``): candidate expansion with the
Spec F decode fused in, ONE pass of the library's top-k / NMS kernels over all (box, class) candidates with
``ids`` = class and ``force_suppress = False`` (replaces the reference's Python loop over classes), and the output
gather; nothing synchronises with the host."""
import torch

from ... import _lib as L


def _det_call(boxes, deltas, scores, means, stds, img_shape, scale_factor, score_thr, iou_thr, delta, max_num):
    L.require_cuda(boxes, deltas, scores)
    n, C = scores.shape
    dev = scores.device
    m = n * (C - 1)
    k = min(m, L.MXD_SORT_CAP)
    cap = min(k, max_num) if max_num > 0 else k
    if m == 0 or cap == 0:
        return (torch.zeros((cap, 5), dtype=torch.float32, device=dev), torch.full((cap,), -1, dtype=torch.int32, device=dev),
                torch.zeros(1, dtype=torch.int32, device=dev))
    dets = torch.empty((cap, 5), dtype=torch.float32, device=dev)       # the library writes every row (zeros / -1 past num)
    labels = torch.empty((cap,), dtype=torch.int32, device=dev)
    num = torch.empty(1, dtype=torch.int32, device=dev)
    ws = L.workspace(L.lib.mxd_det_bboxes_workspace_bytes(n, C, int(max_num)), dev, "det")
    ih, iw = (int(img_shape[0]), int(img_shape[1])) if img_shape is not None else (0, 0)
    L.call("mxd_det_bboxes", L.dl(boxes.float().contiguous()), L.dl(None if deltas is None else deltas.float().contiguous()),
           L.dl(scores.float().contiguous()), L.float4(means), L.float4(stds), ih, iw, 16 / 1000, float(scale_factor),
           float(score_thr), float(iou_thr), float(delta), int(max_num), L.dl(dets), L.dl(labels), L.dl(num), ws.data_ptr(),
           ws.numel(), L.current_stream(dev))
    return dets, labels, num


def multiclass_nms(multi_bboxes, multi_scores, score_thr, iou_thr, max_num=-1, delta=1.0):
    """multi_bboxes (n,4) or (n,4*C); multi_scores (n,C) with column 0 = background.
    Returns (dets (cap,5) [x1,y1,x2,y2,score], labels (cap) i32 = class-1, num (1) i32); rows beyond num are 0 / -1.
    Order: score descending (ties -> lower candidate index (box-major, class-minor)).  At most 8192 candidates
    above the threshold enter the NMS (the in-CTA sort capacity), best first."""
    return _det_call(multi_bboxes, None, multi_scores, (0, 0, 0, 0), (1, 1, 1, 1), None, 1.0, score_thr, iou_thr, delta, max_num)


def get_det_bboxes(rois, cls_score, bbox_pred, img_shape, scale_factor=1.0, score_thr=0.05, iou_thr=0.5, max_per_img=100,
                   target_means=(0, 0, 0, 0), target_stds=(0.1, 0.1, 0.2, 0.2), reg_class_agnostic=False):
    """BBoxHead.get_det_bboxes: rois (n,5), cls_score (n,C) ALREADY softmaxed, bbox_pred (n,4) or (n,4*C) or None."""
    boxes = rois[:, 1:]
    if bbox_pred is None:      # rois are the final boxes: no clip, but still rescaled to the original image frame
        return _det_call(boxes, None, cls_score, target_means, target_stds, None, scale_factor, score_thr, iou_thr, 1.0,
                         max_per_img)
    return _det_call(boxes, bbox_pred, cls_score, target_means, target_stds, img_shape, scale_factor, score_thr, iou_thr, 1.0,
                     max_per_img)
