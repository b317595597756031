"""Per-class score threshold -> class-aware NMS -> top max_num (mmdet-0.5 ``multiclass_nms`` /
``BBoxHead.get_det_bboxes``).  One pass of the library's NMS kernels over all
(box, class) candidates with ``ids`` = class and ``force_suppress = False`` replaces the reference's Python loop
over classes; nothing synchronises with the host."""
import torch

from ... import _lib as L
from ...core.bbox.transforms import delta2bbox
from ...ops.nms import nms_indices


def multiclass_nms(multi_bboxes, multi_scores, score_thr, iou_thr, max_num=-1, delta=1.0):
    """multi_bboxes (n,4) or (n,4*C); multi_scores (n,C) with column 0 = background.
    Returns (dets (cap,5) [x1,y1,x2,y2,score], labels (cap) i32 = class-1, num (1) i32); rows beyond num are 0 / -1.
    Order: score descending (ties -> lower candidate index (box-major, class-minor)).  At most 8192 candidates
    above the threshold enter the NMS (the in-CTA sort capacity); more is a capacity error in the reference's
    own practical range (score_thr 0.05, 1000 proposals)."""
    L.require_cuda(multi_bboxes, multi_scores)
    n, C = multi_scores.shape
    if multi_bboxes.shape[1] == 4:
        boxes = multi_bboxes[:, None, :].expand(n, C - 1, 4)
    else:
        boxes = multi_bboxes.reshape(n, C, 4)[:, 1:, :]
    boxes = boxes.reshape(-1, 4).contiguous().float()
    scores = multi_scores[:, 1:].reshape(-1).contiguous().float()
    ids = torch.arange(C - 1, device=scores.device, dtype=torch.int32).repeat(n)
    m = boxes.shape[0]
    topk = min(m, L.MXD_SORT_CAP)
    cap = min(topk, max_num) if max_num > 0 else topk
    keep, num = nms_indices(boxes, scores, iou_thr, delta=delta, topk=topk, valid_thresh=float(score_thr), ids=ids,
                            force_suppress=False, max_out=cap)
    ok = keep >= 0
    safe = torch.where(ok, keep, torch.zeros_like(keep)).long()
    dets = torch.cat([boxes[safe], scores[safe, None]], 1) * ok[:, None].float()
    labels = torch.where(ok, ids[safe], torch.full_like(keep, -1))
    return dets, labels, num


def get_det_bboxes(rois, cls_score, bbox_pred, img_shape, scale_factor=1.0, score_thr=0.05, iou_thr=0.5, max_per_img=100,
                   target_means=(0, 0, 0, 0), target_stds=(0.1, 0.1, 0.2, 0.2), reg_class_agnostic=False):
    """BBoxHead.get_det_bboxes: rois (n,5), cls_score (n,C) ALREADY softmaxed, bbox_pred (n,4) or (n,4*C)."""
    L.require_cuda(rois, cls_score, bbox_pred)
    n, C = cls_score.shape
    if bbox_pred is None:
        bboxes = rois[:, 1:].contiguous()
    elif reg_class_agnostic or bbox_pred.shape[1] == 4:
        bboxes = delta2bbox(rois[:, 1:].contiguous(), bbox_pred.contiguous(), target_means, target_stds, img_shape)
    else:
        r = rois[:, None, 1:].expand(n, C, 4).reshape(-1, 4).contiguous()
        bboxes = delta2bbox(r, bbox_pred.reshape(-1, 4).contiguous(), target_means, target_stds, img_shape).reshape(n, 4 * C)
    if scale_factor != 1.0:
        bboxes = bboxes / float(scale_factor)
    return multiclass_nms(bboxes, cls_score, score_thr, iou_thr, max_per_img)
