"""Oracle: IoU matrix and max-IoU assigner (SURVEY.md 8(a) Specs D, E).

TEST INFRASTRUCTURE - see oracle/__init__.py.  PARITY UNPINNED by the
reference; module roles = mxdetection/core/bbox and core/anchor
(/root/reference/README.md:16-17).  Pinned by KAT-4 (hand-built tie case)
and torchvision box_iou (delta=0).
"""
import numpy as np

F = np.float32


def bbox_overlaps(b1, b2, delta=1.0):
    """Pairwise IoU (G,4)x(N,4) -> (G,N) f32, Spec D op order."""
    b1 = np.asarray(b1, dtype=F).reshape(-1, 4); b2 = np.asarray(b2, dtype=F).reshape(-1, 4)
    d = F(delta)
    a1 = (((b1[:, 2] - b1[:, 0]) + d) * ((b1[:, 3] - b1[:, 1]) + d)).astype(F)
    a2 = (((b2[:, 2] - b2[:, 0]) + d) * ((b2[:, 3] - b2[:, 1]) + d)).astype(F)
    iw = ((np.minimum(b1[:, None, 2], b2[None, :, 2]) - np.maximum(b1[:, None, 0], b2[None, :, 0])) + d).astype(F)
    ih = ((np.minimum(b1[:, None, 3], b2[None, :, 3]) - np.maximum(b1[:, None, 1], b2[None, :, 1])) + d).astype(F)
    inter = np.where((iw > 0) & (ih > 0), (iw * ih).astype(F), F(0)).astype(F)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (inter / ((a1[:, None] + a2[None, :]) - inter).astype(F)).astype(F)


def max_iou_assign(anchors, gts, gt_labels=None, pos_iou_thr=0.7, neg_iou_thr=0.3,
                   min_pos_iou=0.3, flags=None, delta=1.0):
    """Spec E.  Returns (assigned_gt_inds i32, max_overlaps f32, labels i32), all (N).

    Rows with flags==0 are excluded before assignment and come back as
    assigned=-1, max_overlaps=0, label=0 (the 'unmap' fill)."""
    anchors = np.asarray(anchors, dtype=F).reshape(-1, 4)
    gts = np.asarray(gts, dtype=F).reshape(-1, 4)
    N = anchors.shape[0]; G = gts.shape[0]
    assigned_full = np.full(N, -1, np.int32)
    maxov_full = np.zeros(N, F)
    labels_full = np.zeros(N, np.int32)
    sel = np.ones(N, bool) if flags is None else np.asarray(flags).astype(bool)
    a = anchors[sel]
    n = a.shape[0]
    assigned = np.full(n, -1, np.int32)
    max_ov = np.zeros(n, F)
    if G == 0:
        assigned[:] = 0
    elif n > 0:
        ov = bbox_overlaps(gts, a, delta)          # (G,n)
        argmax = np.argmax(ov, axis=0)              # first max -> lowest g
        max_ov = ov[argmax, np.arange(n)]
        gt_max = ov.max(axis=1)
        neg = (max_ov >= F(0)) & (max_ov < F(neg_iou_thr))
        assigned[neg] = 0
        pos = max_ov >= F(pos_iou_thr)
        assigned[pos] = argmax[pos] + 1
        for g in range(G):
            if gt_max[g] >= F(min_pos_iou):
                assigned[ov[g] == gt_max[g]] = g + 1
    labels = np.zeros(n, np.int32)
    if gt_labels is not None and G > 0:
        gl = np.asarray(gt_labels, np.int32)
        p = assigned > 0
        labels[p] = gl[assigned[p] - 1]
    assigned_full[sel] = assigned
    maxov_full[sel] = max_ov
    labels_full[sel] = labels
    return assigned_full, maxov_full, labels_full
