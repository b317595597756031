"""Anchor flags and anchor -> GT assignment (mmdet-0.5 anchor_target leaf
functions; mxdetection/core/anchor, /root/reference/README.md:16).
Sampling / target packing are native too: mxd_random_sample / mxd_pack_targets."""
import torch

from ... import _lib as L
from ..bbox.assignment import MaxIoUAssigner


def anchor_inside_flags(flat_anchors, valid_flags, img_shape, allowed_border=0):
    """valid & anchor inside the image grown by allowed_border (allowed_border < 0: valid only)."""
    L.require_cuda(flat_anchors, valid_flags)
    h, w = int(img_shape[0]), int(img_shape[1])
    out = torch.empty((flat_anchors.shape[0],), dtype=torch.uint8, device=flat_anchors.device)
    L.call("mxd_inside_flags", L.dl(flat_anchors.contiguous()),
           L.dl(None if valid_flags is None else valid_flags.contiguous()), h, w, float(allowed_border), L.dl(out),
           L.current_stream(out.device))
    return out


def anchor_assign(flat_anchors, inside_flags, gt_bboxes, num_gts=None, gt_labels=None, pos_iou_thr=0.7,
                  neg_iou_thr=0.3, min_pos_iou=0.3):
    """Batched RPN assignment: anchors (N,4) shared by the batch, gts (B,G,4) padded, num_gts (B)."""
    assigner = MaxIoUAssigner(pos_iou_thr, neg_iou_thr, min_pos_iou)
    return assigner.assign_batch(flat_anchors, gt_bboxes, num_gts, gt_labels, flags=inside_flags)
