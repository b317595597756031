"""CPU tests: INDEPENDENT cross-checks of the oracle rows that have no compiled twin in torchvision's C++ kernels
(D2 assigner, F codec, H1 proposal pipeline, N1 sampling were verified only against a restatement
written by the same hand as the kernels).  Every check here goes a structurally different route:

  * D2: a scalar, loop-per-pair Python assigner (no vectorised max / argmax, low-quality rule as a second explicit
        pass) and torchvision's ``Matcher`` mapped through its one documented difference (low-quality matches
        restore the ARGMAX there, assign g here);
  * F : torchvision's ``BoxCoder`` (delta = 0 convention) mapped onto the delta = 1 convention by shifting x2, y2 by +1;
  * H1: a pipeline built from ``torch.sort(stable=True)``, the mapped ``BoxCoder.decode`` and
        ``torchvision.ops.nms`` (delta = 0 on +1-shifted boxes) on integer-valued inputs, where both routes are exact;
  * N1: a sort-based selection.

The oracle itself stays "unpinned by the reference" (/root/reference holds no vectors); these tests remove the
single-author risk for the rows above."""
import math

import numpy as np
import pytest
import torch
import torchvision
from torchvision.models.detection._utils import BoxCoder, Matcher

import oracle

F = np.float32


def rand_boxes(rng, n, h, w, integer=False):
    x1 = rng.uniform(0, w - 8, n); y1 = rng.uniform(0, h - 8, n)
    bw = np.exp(rng.uniform(np.log(4), np.log(w / 2), n)); bh = np.exp(rng.uniform(np.log(4), np.log(h / 2), n))
    b = np.stack([x1, y1, np.minimum(x1 + bw, w - 1), np.minimum(y1 + bh, h - 1)], 1)
    return (np.round(b) if integer else b).astype(F)


# ------------------------------------------------------------------- D2 assigner -----
def scalar_assigner(anchors, gts, pos, neg, min_pos, delta=1.0):
    """Spec D/E with Python loops and np.float32 scalars only."""
    d = F(delta)
    N_, G = len(anchors), len(gts)
    ov = [[F(0)] * N_ for _ in range(G)]
    for g in range(G):
        gx1, gy1, gx2, gy2 = [F(v) for v in gts[g]]
        ag = F(F(F(gx2 - gx1) + d) * F(F(gy2 - gy1) + d))
        for n in range(N_):
            ax1, ay1, ax2, ay2 = [F(v) for v in anchors[n]]
            aa = F(F(F(ax2 - ax1) + d) * F(F(ay2 - ay1) + d))
            iw = F(F(min(ax2, gx2) - max(ax1, gx1)) + d); ih = F(F(min(ay2, gy2) - max(ay1, gy1)) + d)
            inter = F(iw * ih) if (iw > 0 and ih > 0) else F(0)
            ov[g][n] = F(inter / F(F(ag + aa) - inter))
    assigned = [-1] * N_
    maxov = [F(0)] * N_
    for n in range(N_):
        best, arg = ov[0][n], 0
        for g in range(1, G):
            if ov[g][n] > best:           # strict: ties keep the lowest g
                best, arg = ov[g][n], g
        maxov[n] = best
        if F(0) <= best < F(neg):
            assigned[n] = 0
        if best >= F(pos):
            assigned[n] = arg + 1
    for g in range(G):                     # low-quality rule, ascending g: the later g overrides
        gmax = max(ov[g])
        if gmax >= F(min_pos):
            for n in range(N_):
                if ov[g][n] == gmax:
                    assigned[n] = g + 1
    return np.array(assigned, np.int32), np.array(maxov, F)


@pytest.mark.parametrize("thr", [(0.7, 0.3, 0.3), (0.5, 0.5, 0.5), (0.5, 0.4, 0.0)])
def test_assigner_against_scalar_restatement(thr):
    rng = np.random.default_rng(int(thr[0] * 10))
    gts = rand_boxes(rng, 9, 200, 300, integer=True)
    anchors = rand_boxes(rng, 260, 200, 300, integer=True)
    anchors[:9] = gts                      # IoU-1 matches
    anchors[9:12] = anchors[12:15]         # duplicate anchors: tie for a GT's maximum
    a, m, _ = oracle.max_iou_assign(anchors, gts, None, *thr)
    ra, rm = scalar_assigner(anchors, gts, *thr)
    assert np.array_equal(a, ra) and np.array_equal(m, rm)


def test_assigner_against_torchvision_matcher():
    """Matcher(high, low, allow_low_quality_matches=True) on the SAME IoU matrix: identical to Spec E except that its
    low-quality rule restores the anchor's own argmax instead of assigning the claiming GT - so the SETS of positive /
    ignored / negative anchors agree, and the matched GT agrees wherever the anchor's argmax is the claiming GT."""
    rng = np.random.default_rng(3)
    gts = rand_boxes(rng, 12, 300, 400)
    anchors = rand_boxes(rng, 2000, 300, 400)
    anchors[:600] = gts[rng.integers(0, 12, 600)] + rng.normal(0, 2.0, (600, 4)).astype(F)      # jittered copies: real matches
    ov = oracle.bbox_overlaps(gts, anchors, 1.0)
    a, m, _ = oracle.max_iou_assign(anchors, gts, None, 0.7, 0.3, 0.0)       # min_pos 0: Matcher has no such floor
    tv = Matcher(0.7, 0.3, allow_low_quality_matches=True)(torch.from_numpy(ov)).numpy()
    assert np.array_equal(tv >= 0, a > 0)                     # positives (threshold or low-quality)
    assert np.array_equal(tv == Matcher.BELOW_LOW_THRESHOLD, a == 0)
    assert np.array_equal(tv == Matcher.BETWEEN_THRESHOLDS, a == -1)
    same = (a > 0) & (a - 1 == ov.argmax(0))
    assert np.array_equal(tv[same], a[same] - 1) and same.sum() > 20


# ---------------------------------------------------------------------- F codec ------
def test_codec_against_torchvision_boxcoder():
    rng = np.random.default_rng(5)
    p = rand_boxes(rng, 3000, 800, 1344); g = rand_boxes(rng, 3000, 800, 1344)
    shift = np.array([0, 0, 1, 1], F)
    for means, stds in (((0, 0, 0, 0), (1, 1, 1, 1)), ((0, 0, 0, 0), (0.1, 0.1, 0.2, 0.2))):
        coder = BoxCoder(tuple(1.0 / s for s in stds))
        enc = oracle.bbox2delta(p, g, means, stds)
        tv = coder.encode_single(torch.from_numpy(g + shift), torch.from_numpy(p + shift)).numpy()
        assert np.abs(enc - tv).max() <= 2e-5 * max(1.0, np.abs(tv).max())
        d = rng.normal(0, 0.5, (3000, 4)).astype(F)
        d[:, 2:] = np.clip(d[:, 2:] * np.array(stds[2:], F), -4.0, 4.0) / np.array(stds[2:], F)      # inside both clamps
        dec = oracle.delta2bbox(p, d, means, stds, None)
        tvd = coder.decode_single(torch.from_numpy(d), torch.from_numpy(p + shift)).numpy() - shift
        assert np.abs(dec - tvd).max() <= 1e-3              # px; fp32 exp / reassociation on boxes up to 1344 px * e^4
        sel = np.abs(d[:, 2:]).max(1) < 1.0                 # ordinary deltas: the 1e-5-relative bar of Spec F
        assert np.abs(dec[sel] - tvd[sel]).max() <= 1e-5 * 4000


# ------------------------------------------------------------- H1 proposal pipeline --
def test_rpn_pipeline_against_torch_sort_boxcoder_nms():
    """Integer anchors, dyadic deltas with dw = dh = 0: decode is exact in fp32 on both routes, so the independently
    built pipeline must give the SAME proposals, in the same order."""
    rng = np.random.default_rng(9)
    feat_shapes = [(12, 16), (6, 8)]
    strides = [8, 16]
    base = [oracle.gen_base_anchors(s, [8], [0.5, 1, 2]) for s in strides]
    img = (96, 128)
    B = 2
    scores, deltas = [], []
    for (fh, fw) in feat_shapes:
        n = fh * fw * 3
        scores.append((np.round(rng.uniform(0, 1, (B, n)) * 64) / 64).astype(F))           # ties
        dl = np.zeros((B, n, 4), F)
        dl[..., :2] = rng.integers(-4, 5, (B, n, 2)) / 16.0                                # dyadic shifts, dw = dh = 0
        deltas.append(dl)
    cfg = dict(nms_pre=200, nms_post=60, max_num=80, nms_thr=0.7)
    ref, refn = oracle.rpn_proposals(scores, deltas, base, feat_shapes, strides, np.array([img] * B, np.int32), **cfg)
    coder = BoxCoder((1.0, 1.0, 1.0, 1.0))
    shift = torch.tensor([0, 0, 1, 1], dtype=torch.float32)
    for b in range(B):
        cat = []
        for l, (fh, fw) in enumerate(feat_shapes):
            anc = torch.from_numpy(oracle.grid_anchors(base[l], fh, fw, strides[l]))
            s = torch.from_numpy(scores[l][b])
            order = torch.sort(s, descending=True, stable=True).indices[: cfg["nms_pre"]]
            box = coder.decode_single(torch.from_numpy(deltas[l][b])[order], anc[order] + shift) - shift
            box[:, 0::2] = box[:, 0::2].clamp(0, img[1] - 1); box[:, 1::2] = box[:, 1::2].clamp(0, img[0] - 1)
            keep = torchvision.ops.nms(box + shift, s[order], cfg["nms_thr"])[: cfg["nms_post"]]       # delta 1 == delta 0 on x2+1
            cat.append(torch.cat([box[keep], s[order][keep, None]], 1))
        cat = torch.cat(cat)
        if len(cat) > cfg["max_num"]:
            cat = cat[torch.sort(cat[:, 4], descending=True, stable=True).indices[: cfg["max_num"]]]
        assert refn[b] == len(cat)
        assert np.array_equal(ref[b, : len(cat)], cat.numpy()), "image %d" % b


# ---------------------------------------------------------------------- N1 sampling --
def test_random_sample_against_sort_selection():
    rng = np.random.default_rng(13)
    assigned = rng.choice([-1, 0, 0, 0, 1, 2, 3], 5000).astype(np.int32)
    keys = (np.round(rng.random(5000) * 1024) / 1024).astype(F)           # ties: the lower index wins
    for num, frac, ub in ((256, 0.5, -1), (100, 0.7, -1), (512, 0.25, 3), (64, 0.9, 0)):
        pos, neg = oracle.targets.random_sample(assigned, keys, num, frac, ub)
        idx = np.arange(5000)
        order = np.lexsort((idx, -keys.astype(np.float64)))               # key descending, index ascending
        pc = [i for i in order if assigned[i] > 0][: int(num * frac)]
        nneg = num - len(pc)
        if ub >= 0:
            nneg = min(nneg, ub * max(1, len(pc)))
        nc = [i for i in order if assigned[i] == 0][:nneg]
        assert np.array_equal(pos[pos >= 0], np.array(pc, np.int64)) and np.array_equal(neg[neg >= 0], np.array(nc, np.int64))
