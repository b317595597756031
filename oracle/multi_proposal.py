"""TEST INFRASTRUCTURE - CPU restatement of mx.nd.contrib.MultiProposal (mxnet 1.3.0 multi_proposal.cc),
SURVEY.md 8(a) Spec H alt-mode + Spec F MX13 variant.  Parity unpinned by the reference (no source / vectors
mounted); anchored on the classic 9-anchor table (KAT-2) and hand-checked cases in tests/test_oracle.py.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this package."""
import numpy as np

from .anchors import generate_anchors_mx
from .bbox_codec import exp_cr
from .nms import nms

F = np.float32


def multi_proposal(cls_prob, bbox_pred, im_info, rpn_pre_nms_top_n=6000, rpn_post_nms_top_n=300, threshold=0.7,
                   rpn_min_size=16, scales=(4, 8, 16, 32), ratios=(0.5, 1, 2), feature_stride=16):
    """-> rois (N*post_n,5) f32, scores (N*post_n,1) f32.  Strict fp32, operations in multi_proposal.cc order."""
    cls_prob = np.asarray(cls_prob, F); bbox_pred = np.asarray(bbox_pred, F); im_info = np.asarray(im_info, F)
    N, twoA, H, W = cls_prob.shape
    A = twoA // 2
    base = generate_anchors_mx(feature_stride, scales, ratios)
    assert base.shape[0] == A
    s = F(feature_stride)
    post_n = int(rpn_post_nms_top_n)
    rois = np.zeros((N * post_n, 5), F); out_scores = np.zeros((N * post_n, 1), F)
    hh, ww, aa = np.meshgrid(np.arange(H), np.arange(W), np.arange(A), indexing="ij")      # index = (h*W + w)*A + a
    hh = hh.reshape(-1); ww = ww.reshape(-1); aa = aa.reshape(-1)
    for b in range(N):
        im_h, im_w, im_s = im_info[b]
        sx = (ww.astype(F) * s).astype(F); sy = (hh.astype(F) * s).astype(F)
        x1 = (base[aa, 0] + sx).astype(F); y1 = (base[aa, 1] + sy).astype(F)
        x2 = (base[aa, 2] + sx).astype(F); y2 = (base[aa, 3] + sy).astype(F)
        score = cls_prob[b, A + aa, hh, ww].astype(F).copy()
        dx = bbox_pred[b, 4 * aa + 0, hh, ww]; dy = bbox_pred[b, 4 * aa + 1, hh, ww]
        dw = bbox_pred[b, 4 * aa + 2, hh, ww]; dh = bbox_pred[b, 4 * aa + 3, hh, ww]
        bw = ((x2 - x1).astype(F) + F(1)).astype(F); bh = ((y2 - y1).astype(F) + F(1)).astype(F)
        cx = (x1 + (F(0.5) * (bw - F(1)).astype(F)).astype(F)).astype(F)
        cy = (y1 + (F(0.5) * (bh - F(1)).astype(F)).astype(F)).astype(F)
        pcx = ((dx * bw).astype(F) + cx).astype(F); pcy = ((dy * bh).astype(F) + cy).astype(F)
        pw = (exp_cr(dw) * bw).astype(F); ph = (exp_cr(dh) * bh).astype(F)
        hw = (F(0.5) * (pw - F(1)).astype(F)).astype(F); hhh = (F(0.5) * (ph - F(1)).astype(F)).astype(F)
        bx1 = np.maximum(np.minimum((pcx - hw).astype(F), F(im_w - F(1))), F(0))
        by1 = np.maximum(np.minimum((pcy - hhh).astype(F), F(im_h - F(1))), F(0))
        bx2 = np.maximum(np.minimum((pcx + hw).astype(F), F(im_w - F(1))), F(0))
        by2 = np.maximum(np.minimum((pcy + hhh).astype(F), F(im_h - F(1))), F(0))
        real_h = int(F(im_h) / s); real_w = int(F(im_w) / s)
        score[(hh >= real_h) | (ww >= real_w)] = F(-1)
        ms = F(F(rpn_min_size) * im_s)
        iw = ((bx2 - bx1).astype(F) + F(1)).astype(F); ih = ((by2 - by1).astype(F) + F(1)).astype(F)
        small = (iw < ms) | (ih < ms)
        g = F(ms * F(0.5))
        bx1 = np.where(small, (bx1 - g).astype(F), bx1); by1 = np.where(small, (by1 - g).astype(F), by1)
        bx2 = np.where(small, (bx2 + g).astype(F), bx2); by2 = np.where(small, (by2 + g).astype(F), by2)
        score[small] = F(-1)
        boxes = np.stack([bx1, by1, bx2, by2], 1).astype(F)
        keep = nms(boxes, score, threshold, delta=1.0, topk=rpn_pre_nms_top_n, max_out=post_n)
        nk = len(keep)
        for r in range(post_n):
            i = keep[r % nk]
            rois[b * post_n + r] = [b, *boxes[i]]
            out_scores[b * post_n + r, 0] = score[i]
    return rois, out_scores
