"""Max-IoU assigner (SURVEY.md 8(a) Spec E; mmdet-0.5 bbox_assign_wrt_overlaps /
MaxIoUAssigner role of mxdetection/core/bbox, /root/reference/README.md:17).

The G x N overlap matrix is never materialised: a fused two-pass kernel
produces assigned_gt_inds / max_overlaps / labels directly."""
import torch

from ... import _lib as L


class AssignResult:
    def __init__(self, num_gts, gt_inds, max_overlaps, labels=None):
        self.num_gts = num_gts
        self.gt_inds = gt_inds            # int32: -1 ignore, 0 negative, g+1 positive
        self.max_overlaps = max_overlaps
        self.labels = labels


class MaxIoUAssigner:
    def __init__(self, pos_iou_thr, neg_iou_thr, min_pos_iou=0.0, delta=1.0):
        self.pos_iou_thr = float(pos_iou_thr)
        self.neg_iou_thr = float(neg_iou_thr)
        self.min_pos_iou = float(min_pos_iou)
        self.delta = float(delta)

    def assign_batch(self, bboxes, gt_bboxes, num_gts=None, gt_labels=None, flags=None):
        """bboxes (N,4) shared by the batch; gt_bboxes (B,G,4) padded; num_gts (B) i32; gt_labels (B,G) i32;
        flags (N) or (B,N) u8.  Returns (assigned (B,N) i32, max_overlaps (B,N) f32, labels (B,N) i32)."""
        L.require_cuda(bboxes, gt_bboxes, num_gts, gt_labels, flags)
        bboxes = bboxes.contiguous(); gt_bboxes = gt_bboxes.contiguous()
        squeeze = gt_bboxes.dim() == 2
        B = 1 if squeeze else gt_bboxes.shape[0]
        G = gt_bboxes.shape[-2]
        N = bboxes.shape[0]
        dev = bboxes.device
        assigned = torch.empty((B, N), dtype=torch.int32, device=dev)
        max_ov = torch.empty((B, N), dtype=torch.float32, device=dev)
        labels = torch.empty((B, N), dtype=torch.int32, device=dev)
        nbytes = L.lib.mxd_max_iou_assign_workspace_bytes(B, G)
        ws = L.workspace(nbytes, dev, "assign")
        L.call("mxd_max_iou_assign", L.dl(bboxes), L.dl(gt_bboxes),
               L.dl(None if num_gts is None else num_gts.contiguous()),
               L.dl(None if gt_labels is None else gt_labels.contiguous()),
               L.dl(None if flags is None else flags.contiguous()),
               L.dl(assigned), L.dl(max_ov), L.dl(labels), self.pos_iou_thr, self.neg_iou_thr, self.min_pos_iou,
               self.delta, ws.data_ptr(), ws.numel(), L.current_stream(dev))
        if squeeze:
            return assigned[0], max_ov[0], labels[0]
        return assigned, max_ov, labels

    def assign(self, bboxes, gt_bboxes, gt_labels=None):
        """Single image, mmdet signature -> AssignResult."""
        a, m, l = self.assign_batch(bboxes, gt_bboxes.reshape(-1, 4), None, gt_labels)
        return AssignResult(gt_bboxes.shape[0], a, m, l if gt_labels is not None else None)


def bbox_assign(proposals, gt_bboxes, gt_labels=None, pos_iou_thr=0.5, neg_iou_thr=0.5, min_pos_iou=0.5):
    """mmdet-0.5 functional form -> (assigned_gt_inds, argmax-free max_overlaps, labels)."""
    return MaxIoUAssigner(pos_iou_thr, neg_iou_thr, min_pos_iou).assign_batch(
        proposals, gt_bboxes.reshape(-1, 4), None, gt_labels)
