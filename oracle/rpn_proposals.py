"""Oracle: RPN proposal stage (SURVEY.md 8(a) Spec H, row H1).

TEST INFRASTRUCTURE - see oracle/__init__.py.  PARITY UNPINNED by the
reference; module role = mxdetection/models/rpn_heads
(/root/reference/README.md:28).
"""
import numpy as np

from .anchors import grid_anchors
from .bbox_codec import delta2bbox
from .nms import nms, stable_order_desc

F = np.float32


def topk_stable(scores, k):
    """Indices of the top-k by (score desc, index asc), sorted the same way."""
    order = stable_order_desc(scores)
    return order[:k].astype(np.int32) if (k > 0 and len(order) > k) else order.astype(np.int32)


def rpn_proposals_single(scores_lvls, deltas_lvls, base_anchors_lvls, feat_shapes, strides,
                         img_shape, nms_pre=2000, nms_thr=0.7, nms_post=1000, max_num=1000,
                         min_bbox_size=0, means=(0, 0, 0, 0), stds=(1, 1, 1, 1), delta=1.0,
                         return_stages=False):
    """One image.  scores_lvls[l]: (H*W*A) activated, (y,x,a) order; deltas_lvls[l]: (H*W*A,4).

    Returns (proposals (max_num,5) zero padded, num_valid).  With
    return_stages also a per-level list of dict(idx, boxes, valid, keep) for
    the stage-wise parity tests."""
    props = []
    stages = []
    for l, (s, d) in enumerate(zip(scores_lvls, deltas_lvls)):
        s = np.asarray(s, dtype=F).reshape(-1); d = np.asarray(d, dtype=F).reshape(-1, 4)
        H, W = feat_shapes[l]
        idx = topk_stable(s, nms_pre)
        anchors = grid_anchors(base_anchors_lvls[l], H, W, strides[l])[idx]
        boxes = delta2bbox(anchors, d[idx], means, stds, max_shape=img_shape)
        sc = s[idx]
        if min_bbox_size > 0:
            w = ((boxes[:, 2] - boxes[:, 0]) + F(1)).astype(F); h = ((boxes[:, 3] - boxes[:, 1]) + F(1)).astype(F)
            valid = (w >= F(min_bbox_size)) & (h >= F(min_bbox_size))
        else:
            valid = np.ones(len(idx), bool)
        keep = nms(boxes, sc, nms_thr, delta=delta, valid_mask=valid, max_out=nms_post)
        keep = keep[:nms_post] if nms_post > 0 else keep
        props.append(np.concatenate([boxes[keep], sc[keep, None]], axis=1))
        stages.append(dict(idx=idx, boxes=boxes, scores=sc, valid=valid, keep=keep))
    cat = np.concatenate(props, axis=0) if props else np.zeros((0, 5), F)
    if len(cat) > max_num:
        cat = cat[stable_order_desc(cat[:, 4])[:max_num]]
    out = np.zeros((max_num, 5), dtype=F)
    out[:len(cat)] = cat
    if return_stages:
        return out, len(cat), stages
    return out, len(cat)


def rpn_proposals(scores, deltas, base_anchors_lvls, feat_shapes, strides, img_shapes, **cfg):
    """Batch: scores[l] (B, H*W*A), deltas[l] (B, H*W*A, 4); img_shapes (B,2)=(h,w)."""
    B = scores[0].shape[0]
    outs, nums = [], []
    for b in range(B):
        o, n = rpn_proposals_single([s[b] for s in scores], [d[b] for d in deltas],
                                    base_anchors_lvls, feat_shapes, strides,
                                    tuple(int(v) for v in img_shapes[b]), **cfg)
        outs.append(o); nums.append(n)
    return np.stack(outs), np.asarray(nums, np.int32)
