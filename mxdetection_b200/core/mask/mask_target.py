"""Mask targets for the positive RoIs: crop + resize of the matched GT mask to
``mask_size`` x ``mask_size`` = RoIAlign of a 1-channel map, so it runs on the RoIAlign kernels.

Stated contract (the reference's cv2 resize is not reproducible on a device): target = [RoIAlign(mask, roi,
mask_size, spatial_scale 1, sample_ratio 2) >= 0.5], Spec A arithmetic."""
import torch

from ... import _lib as L
from ...ops.roi_align import roi_align_forward


def mask_target(pos_proposals, pos_assigned_gt_inds, gt_masks, mask_size=28, sample_ratio=2, binarize=True):
    """pos_proposals (P,4) image coords; pos_assigned_gt_inds (P) int (0-based); gt_masks (G,H,W) u8/float."""
    L.require_cuda(pos_proposals, pos_assigned_gt_inds, gt_masks)
    data = gt_masks[:, None].float().contiguous()
    rois = torch.cat([pos_assigned_gt_inds.float()[:, None], pos_proposals[:, :4].float()], 1).contiguous()
    out = roi_align_forward(data, rois, (mask_size, mask_size), 1.0, sample_ratio)[:, 0]
    return (out >= 0.5).to(torch.uint8) if binarize else out
