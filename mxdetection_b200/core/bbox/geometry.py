"""Pairwise IoU (SURVEY.md 8(a) Spec D; mmdet-0.5 bbox_overlaps, mx.nd.contrib.box_iou)."""
import torch

from ... import _lib as L


def bbox_overlaps(bboxes1, bboxes2, mode="iou", delta=1.0):
    """(G,4) x (N,4) -> (G,N) f32.  delta=1 (mmdet lineage) or 0 (box_iou)."""
    if mode != "iou":
        raise NotImplementedError("only mode='iou' is on the hot path")
    L.require_cuda(bboxes1, bboxes2)
    out = torch.empty((bboxes1.shape[0], bboxes2.shape[0]), dtype=torch.float32, device=bboxes1.device)
    L.call("mxd_bbox_overlaps", L.dl(bboxes1.contiguous()), L.dl(bboxes2.contiguous()), L.dl(out), float(delta),
           L.current_stream(out.device))
    return out
