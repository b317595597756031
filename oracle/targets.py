"""TEST INFRASTRUCTURE - CPU restatements of the SURVEY.md 8(f) "next" rows: target sampling + packing (N1),
detection post-processing (N2), mask targets (N4).  mmdet-0.5 roles of mxdetection/core/{anchor,bbox,mask} and
models/bbox_heads (/root/reference/README.md:16-18,29); parity unpinned by the reference (nothing mounted)."""
import numpy as np

from .bbox_codec import bbox2delta, delta2bbox
from .nms import nms, stable_order_desc
from .roi_align import roi_align_forward

F = np.float32


def random_sample(assigned, keys, num, pos_fraction, neg_pos_ub=-1):
    """Largest-key-first sampling (ties -> lower index): the device RNG contract of core/bbox/sampling.py."""
    assigned = np.asarray(assigned); keys = np.asarray(keys, F)
    kp = int(num * pos_fraction)
    pos = np.nonzero(assigned > 0)[0]; neg = np.nonzero(assigned == 0)[0]
    pos = pos[stable_order_desc(keys[pos])][:kp]
    quota = num - len(pos)
    if neg_pos_ub >= 0:
        quota = min(quota, int(neg_pos_ub * max(1, len(pos))))
    neg = neg[stable_order_desc(keys[neg])][:max(quota, 0)]
    return pos.astype(np.int32), neg.astype(np.int32)


def pack_targets(anchors, assigned, gt_bboxes, pos, neg, gt_labels=None, means=(0, 0, 0, 0), stds=(1, 1, 1, 1),
                 pos_weight=-1.0):
    n = anchors.shape[0]
    labels = np.zeros(n, np.int32); lw = np.zeros(n, F); tgt = np.zeros((n, 4), F); tw = np.zeros((n, 4), F)
    if len(pos):
        g = assigned[pos] - 1
        tgt[pos] = bbox2delta(anchors[pos], gt_bboxes[g], means, stds)
        tw[pos] = 1
        labels[pos] = 1 if gt_labels is None else gt_labels[g]
        lw[pos] = 1.0 if pos_weight <= 0 else pos_weight
    lw[neg] = 1
    return labels, lw, tgt, tw


def multiclass_nms(multi_bboxes, multi_scores, score_thr, iou_thr, max_num=-1, delta=1.0):
    """All (box, class>0) candidates, class-aware greedy NMS in global score order, first max_num."""
    multi_bboxes = np.asarray(multi_bboxes, F); multi_scores = np.asarray(multi_scores, F)
    n, C = multi_scores.shape
    boxes = (np.repeat(multi_bboxes[:, None, :], C - 1, 1) if multi_bboxes.shape[1] == 4
             else multi_bboxes.reshape(n, C, 4)[:, 1:, :]).reshape(-1, 4)
    scores = multi_scores[:, 1:].reshape(-1)
    ids = np.tile(np.arange(C - 1, dtype=np.int32), n)
    keep = nms(boxes, scores, iou_thr, delta=delta, valid_thresh=score_thr, ids=ids, force_suppress=False, max_out=max_num)
    return np.concatenate([boxes[keep], scores[keep, None]], 1).astype(F), ids[keep]


def get_det_bboxes(rois, cls_score, bbox_pred, img_shape, scale_factor=1.0, score_thr=0.05, iou_thr=0.5, max_per_img=100,
                   target_means=(0, 0, 0, 0), target_stds=(0.1, 0.1, 0.2, 0.2)):
    n, C = cls_score.shape
    if bbox_pred is None:          # rois are final (mmdet 0.5: `bboxes = rois[:, 1:]`), still divided by scale_factor
        b = np.asarray(rois, F)[:, 1:].copy()
    elif bbox_pred.shape[1] == 4:
        b = delta2bbox(rois[:, 1:], bbox_pred, target_means, target_stds, img_shape)
    else:
        r = np.repeat(rois[:, None, 1:], C, 1).reshape(-1, 4)
        b = delta2bbox(r, bbox_pred.reshape(-1, 4), target_means, target_stds, img_shape).reshape(n, 4 * C)
    if scale_factor != 1.0:
        b = (b / F(scale_factor)).astype(F)
    return multiclass_nms(b, cls_score, score_thr, iou_thr, max_per_img)


def mask_target(pos_proposals, pos_gt_inds, gt_masks, mask_size=28, sample_ratio=2, thr=0.5):
    data = np.asarray(gt_masks, F)[:, None]
    rois = np.concatenate([np.asarray(pos_gt_inds, F)[:, None], np.asarray(pos_proposals, F)[:, :4]], 1)
    return (roi_align_forward(data, rois, (mask_size, mask_size), 1.0, sample_ratio)[:, 0] >= F(thr)).astype(np.uint8)


def paste_masks(mask_pred, det_bboxes, img_shape, det_labels=None, scale_factor=1.0, thr=0.5):
    """Spec N4 (mask paste; FCNMaskHead.get_seg_masks of mmdet 0.5 with the resize restated): per detection the integer
    box x1i = trunc(x1 / scale), w = max(trunc(x2 / scale) - x1i + 1, 1) (same in y); canvas pixel (x, y) inside it samples the
    S x S probability map at s = (d + 0.5) * (S / ext) - 0.5 (half-pixel convention; s < 0 -> index 0, s >= S-1 -> index S-1,
    fraction 0 at both borders); top = m00 * (1 - lx) + m01 * lx, bot likewise, v = top * (1 - ly) + bot * ly, strict fp32;
    pixel = v > thr.  mask_pred (n,C,S,S) with labels (class c reads channel c + 1) or (n,S,S)."""
    mp = np.asarray(mask_pred, F)
    n = mp.shape[0]
    S = mp.shape[-1]
    H, W = int(img_shape[0]), int(img_shape[1])
    out = np.zeros((n, H, W), np.uint8)
    sc = F(scale_factor)

    def axis(ext):
        d = np.arange(ext, dtype=F)
        s = ((d + F(0.5)) * (F(S) / F(ext))).astype(F) - F(0.5)
        lo = np.floor(s).astype(np.int64)
        fr = (s - lo.astype(F)).astype(F)
        under = lo < 0; over = lo >= S - 1
        lo = np.where(under, 0, np.where(over, S - 1, lo))
        fr = np.where(under | over, F(0), fr).astype(F)
        return lo, np.minimum(lo + 1, S - 1), fr

    for i in range(n):
        bb = (np.asarray(det_bboxes[i, :4], F) / sc).astype(F)
        x1, y1, x2, y2 = [int(v) for v in bb]                      # truncation toward zero
        w = max(x2 - x1 + 1, 1); h = max(y2 - y1 + 1, 1)
        m = mp[i] if mp.ndim == 3 else mp[i, int(det_labels[i]) + 1]
        xa, xb, lx = axis(w); ya, yb, ly = axis(h)
        hx = (F(1) - lx).astype(F); hy = (F(1) - ly).astype(F)
        top = ((m[ya][:, xa] * hx[None, :]).astype(F) + (m[ya][:, xb] * lx[None, :]).astype(F)).astype(F)
        bot = ((m[yb][:, xa] * hx[None, :]).astype(F) + (m[yb][:, xb] * lx[None, :]).astype(F)).astype(F)
        v = ((top * hy[:, None]).astype(F) + (bot * ly[:, None]).astype(F)).astype(F)
        box = (v > F(thr)).astype(np.uint8)
        ys = slice(max(y1, 0), min(y1 + h, H)); xs = slice(max(x1, 0), min(x1 + w, W))
        if ys.stop > ys.start and xs.stop > xs.start:
            out[i, ys, xs] = box[ys.start - y1: ys.stop - y1, xs.start - x1: xs.stop - x1]
    return out
