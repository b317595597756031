"""Mask targets for the positive RoIs and mask paste at test time (mxdetection/core/mask,
/root/reference/README.md:18; ``mask_target`` / ``FCNMaskHead.get_seg_masks`` of mmdet 0.5).

Stated contract (the reference's cv2 resize is not reproducible on a device): target = [RoIAlign(mask, roi,
mask_size, spatial_scale 1, sample_ratio 2) >= 0.5] with Spec A arithmetic in strict fp32 on the uint8 mask bytes
(``mxd_mask_target``: no fp32 copy of the mask stack, bit-exact against the CPU restatement); paste = half-pixel bilinear
resize of the S x S probabilities to the integer detection box, > thr (``mxd_paste_masks``, Spec N4)."""
import torch

from ... import _lib as L


def mask_target(pos_proposals, pos_assigned_gt_inds, gt_masks, mask_size=28, sample_ratio=2, binarize=True, thr=0.5):
    """pos_proposals (P,>=4) f32 image coords; pos_assigned_gt_inds (P) int32 (0-based); gt_masks (G,H,W) uint8.
    Returns (P,S,S) uint8 targets (binarize) or the f32 RoIAlign values."""
    L.require_cuda(pos_proposals, pos_assigned_gt_inds, gt_masks)
    if gt_masks.dtype != torch.uint8:
        raise TypeError("gt_masks must be uint8 (the kernel reads the mask bytes directly)")
    if pos_assigned_gt_inds.dtype != torch.int32:
        raise TypeError("pos_assigned_gt_inds must be int32")
    P = pos_proposals.shape[0]
    out = torch.empty((P, mask_size, mask_size), dtype=torch.uint8 if binarize else torch.float32, device=gt_masks.device)
    L.call("mxd_mask_target", L.dl(gt_masks.contiguous()), L.dl(pos_proposals.contiguous()),
           L.dl(pos_assigned_gt_inds.contiguous()), L.dl(out), int(mask_size), int(sample_ratio), float(thr),
           L.current_stream(gt_masks.device))
    return out


def paste_masks(mask_pred, det_bboxes, img_shape, det_labels=None, scale_factor=1.0, thr=0.5):
    """mask_pred (n,C,S,S) or (n,S,S) f32 probabilities; det_bboxes (n,>=4); det_labels (n) int32 or None (class c reads
    channel c+1).  Returns (n,img_h,img_w) uint8 masks in the original image frame (bbox / scale_factor)."""
    L.require_cuda(mask_pred, det_bboxes, det_labels)
    n = mask_pred.shape[0]
    out = torch.empty((n, int(img_shape[0]), int(img_shape[1])), dtype=torch.uint8, device=mask_pred.device)
    L.call("mxd_paste_masks", L.dl(mask_pred.contiguous()), L.dl(None if det_labels is None else det_labels.contiguous()),
           L.dl(det_bboxes.contiguous()), L.dl(out), float(scale_factor), float(thr), L.current_stream(mask_pred.device))
    return out
