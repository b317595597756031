"""mxdetection/models/mask_heads (/root/reference/README.md:30): the box-arithmetic tail of the mask head -
``FCNMaskHead.get_seg_masks`` of mmdet 0.5 (the convolutions are out of scope)."""
from ...core.mask import paste_masks


def get_seg_masks(mask_pred, det_bboxes, det_labels, ori_shape, scale_factor=1.0, rescale=True, thr_binary=0.5):
    """mask_pred (n,C,S,S) f32 ALREADY sigmoided (column 0 = background); det_bboxes (n,>=4); det_labels (n) int32.
    Returns (n,H,W) uint8 masks of the original image (one ``mxd_paste_masks`` call; the reference loops over
    detections on the host with cv2)."""
    return paste_masks(mask_pred, det_bboxes, ori_shape[:2], det_labels, scale_factor if rescale else 1.0, thr_binary)
