"""Oracle: hard NMS (SURVEY.md 8(a) Spec B, mx.nd.contrib.box_nms semantics).

TEST INFRASTRUCTURE - see oracle/__init__.py.  PARITY UNPINNED by the
reference; contract = mxnet 1.3.0 box_nms (/root/reference/README.md:37),
hosted by mxdetection/ops (/root/reference/README.md:24).  Pinned by KAT-1
(the box_nms docstring example) and by torchvision.ops.nms keep lists.

Canonical rule (BASELINE.json north_star): strict `iou > thr`, stable
score-descending order with ties broken by the lower original index.
"""
import numpy as np

F = np.float32


def _areas(boxes, delta):
    d = F(delta)
    w = np.maximum((boxes[:, 2] - boxes[:, 0]) + d, F(0)).astype(F)
    h = np.maximum((boxes[:, 3] - boxes[:, 1]) + d, F(0)).astype(F)
    return (w * h).astype(F)


def box_iou_pair(a, b, delta=0.0):
    """IoU of two boxes in Spec B op order (used by tests)."""
    d = F(delta)
    a = np.asarray(a, F); b = np.asarray(b, F)
    iw = (min(a[2], b[2]) - max(a[0], b[0])) + d
    ih = (min(a[3], b[3]) - max(a[1], b[1])) + d
    if iw <= 0 or ih <= 0:
        return F(0)
    inter = F(iw * ih)
    aa = F(max(F((a[2] - a[0]) + d), F(0)) * max(F((a[3] - a[1]) + d), F(0)))
    ab = F(max(F((b[2] - b[0]) + d), F(0)) * max(F((b[3] - b[1]) + d), F(0)))
    return F(inter / F(F(aa + ab) - inter))


def stable_order_desc(scores):
    """Indices sorted by score DESC, ties -> lower index first."""
    scores = np.asarray(scores, F)
    return np.argsort(-scores.astype(np.float64), kind="stable")


def nms(boxes, scores, iou_thr, delta=0.0, topk=-1, valid_thresh=-np.inf,
        ids=None, force_suppress=True, max_out=-1, valid_mask=None):
    """Greedy NMS.  Returns keep indices (int32, score order).

    valid_mask: optional bool (n) - False rows are dropped before sorting
    (min-size filter of Spec H)."""
    boxes = np.ascontiguousarray(boxes, dtype=F).reshape(-1, 4)
    scores = np.ascontiguousarray(scores, dtype=F).reshape(-1)
    n = boxes.shape[0]
    thr = F(iou_thr)
    d = F(delta)
    order = stable_order_desc(scores)
    sel = scores[order] > F(valid_thresh) if np.isfinite(valid_thresh) else np.ones(n, bool)
    if valid_mask is not None:
        sel = sel & np.asarray(valid_mask, bool)[order]
    order = order[sel]
    if topk > 0:
        order = order[:topk]
    b = boxes[order]
    area = _areas(b, delta)
    m = len(order)
    suppressed = np.zeros(m, dtype=bool)
    keep = []
    for r in range(m):
        if suppressed[r]:
            continue
        keep.append(order[r])
        if max_out > 0 and len(keep) >= max_out:
            break
        rest = slice(r + 1, m)
        iw = ((np.minimum(b[r, 2], b[rest, 2]) - np.maximum(b[r, 0], b[rest, 0])) + d).astype(F)
        ih = ((np.minimum(b[r, 3], b[rest, 3]) - np.maximum(b[r, 1], b[rest, 1])) + d).astype(F)
        pos = (iw > 0) & (ih > 0)
        inter = (iw * ih).astype(F)
        union = ((area[r] + area[rest]) - inter).astype(F)
        with np.errstate(divide="ignore", invalid="ignore"):
            iou = (inter / union).astype(F)
        hit = pos & (iou > thr)
        if ids is not None and not force_suppress:
            idv = np.asarray(ids)[order]
            hit &= (idv[rest] == idv[r])
        suppressed[rest] |= hit
    return np.asarray(keep, dtype=np.int32)


def _to_corner(b):
    x, y, w, h = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    hw = (w / F(2)).astype(F); hh = (h / F(2)).astype(F)
    return np.stack([x - hw, y - hh, x + hw, y + hh], -1).astype(F)


def _to_center(b):
    x1, y1, x2, y2 = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    w = (x2 - x1).astype(F); h = (y2 - y1).astype(F)
    return np.stack([x1 + (w / F(2)).astype(F), y1 + (h / F(2)).astype(F), w, h], -1).astype(F)


def box_nms_mx(data, overlap_thresh=0.5, valid_thresh=0.0, topk=-1, coord_start=2,
               score_index=1, id_index=-1, force_suppress=False,
               in_format="corner", out_format="corner", return_index=False):
    """MXNet tensor API: (..., N, K) in -> same shape out, kept rows in score
    order at the top, every other row filled with -1 (Spec B output B)."""
    data = np.ascontiguousarray(data, dtype=F)
    shape = data.shape
    N, K = shape[-2], shape[-1]
    x = data.reshape(-1, N, K)
    out = np.full_like(x, F(-1))
    index = np.full(x.shape[:2], -1, dtype=np.int32)
    for bi in range(x.shape[0]):
        rows = x[bi]
        boxes = rows[:, coord_start:coord_start + 4]
        if in_format == "center":
            boxes = _to_corner(boxes)
        ids = rows[:, id_index] if id_index >= 0 else None
        keep = nms(boxes, rows[:, score_index], overlap_thresh, delta=0.0, topk=topk,
                   valid_thresh=valid_thresh, ids=ids,
                   force_suppress=force_suppress or id_index < 0)
        k = len(keep)
        out[bi, :k] = rows[keep]
        if in_format != out_format:
            cb = out[bi, :k, coord_start:coord_start + 4]
            out[bi, :k, coord_start:coord_start + 4] = _to_corner(cb) if out_format == "corner" else _to_center(cb)
        index[bi, :k] = keep
    out = out.reshape(shape)
    if return_index:
        return out, index.reshape(shape[:-1])
    return out
