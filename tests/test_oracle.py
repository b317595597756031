"""CPU tests: the NumPy oracle against known-answer vectors
and the committed torchvision golden fixtures; the C restatement against the NumPy oracle."""
import glob
import os

import numpy as np
import pytest

import oracle
from oracle import cref

F = np.float32


# ---------------------------------------------------------------- Spec A ----------
@pytest.mark.parametrize("name", ["roi_align_a", "roi_align_b", "roi_align_c"])
def test_roi_align_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    ps = tuple(int(v) for v in g["pooled"]); sc = float(g["scale"]); sr = int(g["sample_ratio"])
    out = oracle.roi_align_forward(g["data"], g["rois"], ps, sc, sr)
    assert np.array_equal(out, g["out"]), "forward must be bit-exact with the compiled cross-oracle"
    gin = oracle.roi_align_backward(g["grad_out"], g["rois"], g["data"].shape, ps, sc, sr)
    ref = g["grad_in"]
    assert np.all(np.abs(gin - ref) <= 1e-4 * np.maximum(1, np.abs(ref)))   # bwd bar: 1e-4 (accumulation order)
    # C restatement
    assert np.array_equal(cref.roi_align_forward(g["data"], g["rois"], ps, sc, sr), g["out"])
    gc = cref.roi_align_backward(g["grad_out"], g["rois"], g["data"].shape, ps, sc, sr)
    assert np.all(np.abs(gc - ref) <= 1e-4 * np.maximum(1, np.abs(ref)))


def test_roi_align_kat5_constant_and_ramp():
    H, W = 40, 60
    rois = np.array([[0, 20, 16, 150, 120], [0, 40.5, 33.25, 97.75, 80.5]], F)   # interior RoIs
    const = np.full((1, 2, H, W), 3.25, F)
    out = oracle.roi_align_forward(const, rois, (7, 7), 0.25, 2)
    assert np.allclose(out, 3.25, rtol=0, atol=1e-6)
    yy, xx = np.meshgrid(np.arange(H, dtype=F), np.arange(W, dtype=F), indexing="ij")
    ramp = (0.5 * xx + 0.25 * yy)[None, None].astype(F)
    out = oracle.roi_align_forward(ramp, rois, (7, 7), 0.25, 2)
    for r in range(2):
        x1, y1, x2, y2 = rois[r, 1:] * 0.25
        bw = (x2 - x1) / 7; bh = (y2 - y1) / 7
        cx = x1 + (np.arange(7) + 0.5) * bw; cy = y1 + (np.arange(7) + 0.5) * bh
        expect = 0.5 * cx[None, :] + 0.25 * cy[:, None]      # ramp value at the bin centres
        assert np.allclose(out[r, 0], expect, atol=1e-4)


def test_roi_align_edge_cases():
    data = np.random.default_rng(0).standard_normal((2, 3, 10, 12)).astype(F)
    rois = np.array([[-1, 0, 0, 10, 10],        # negative batch -> zeros
                     [0, -500, -500, -400, -400],  # far outside: every sample skipped -> zeros
                     [1, 44, 36, 47.9, 39.9],      # bottom-right corner clamp
                     [0, 5, 5, 5, 5]], F)          # degenerate: width clamps to 1
    out = oracle.roi_align_forward(data, rois, (2, 2), 1.0, 2)
    assert np.all(out[0] == 0) and np.all(out[1] == 0)
    assert np.all(np.isfinite(out))
    assert np.array_equal(out, cref.roi_align_forward(data, rois, (2, 2), 1.0, 2))
    assert oracle.roi_align_forward(data, rois[:0], (2, 2), 1.0, 2).shape == (0, 3, 2, 2)


# ---------------------------------------------------------------- Spec B ----------
def test_box_nms_kat1_docstring():
    x = np.array([[0, .5, .1, .1, .2, .2], [1, .4, .1, .1, .2, .2], [0, .3, .1, .1, .14, .14], [2, .6, .5, .5, .7, .8]], F)
    out = oracle.box_nms_mx(x, overlap_thresh=0.1, coord_start=2, score_index=1, id_index=0, force_suppress=True)
    exp = np.array([[2, .6, .5, .5, .7, .8], [0, .5, .1, .1, .2, .2], [-1] * 6, [-1] * 6], F)
    assert np.array_equal(out, exp)
    # class-aware: box1 (id 1) survives because only id-0 boxes suppress id-0 boxes
    out2 = oracle.box_nms_mx(x, overlap_thresh=0.1, coord_start=2, score_index=1, id_index=0, force_suppress=False)
    assert np.array_equal(out2[:3, 0], np.array([2, 0, 1], F)) and np.all(out2[3] == -1)


@pytest.mark.parametrize("name", ["nms_a", "nms_b", "nms_c"])
def test_nms_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    keep = oracle.nms(g["boxes"], g["scores"], float(g["thr"]), delta=0.0)
    assert np.array_equal(keep, g["keep"])
    assert np.array_equal(cref.nms(g["boxes"], g["scores"], float(g["thr"]), delta=0.0), g["keep"])
    iou = oracle.bbox_overlaps(g["boxes"][:50], g["boxes"][:200], delta=0.0)
    assert np.allclose(iou, g["iou_50x200"], rtol=0, atol=1e-6)


def test_nms_ties_topk_validthresh():
    boxes = np.array([[0, 0, 10, 10], [0, 0, 10, 10], [20, 20, 30, 30], [0, 0, 9, 9], [50, 50, 60, 60]], F)
    scores = np.array([.5, .5, .5, .9, .1], F)
    assert oracle.nms(boxes, scores, 0.5).tolist() == [3, 2, 4]               # 0 and 1 suppressed by 3
    assert oracle.nms(boxes, scores, 0.9).tolist() == [3, 0, 2, 4]            # tie 0/1: lower index first, 1 == 0
    assert oracle.nms(boxes, scores, 0.9, topk=2).tolist() == [3, 0]
    assert oracle.nms(boxes, scores, 0.9, valid_thresh=0.1).tolist() == [3, 0, 2]
    assert oracle.nms(boxes, scores, 0.9, max_out=1).tolist() == [3]
    assert oracle.nms(boxes[:0], scores[:0], 0.5).tolist() == []


# ---------------------------------------------------------------- Spec C ----------
def test_anchors_kat2():
    a = oracle.generate_anchors_mx(16, (8, 16, 32), (0.5, 1, 2))
    assert a.shape == (9, 4)
    assert a[0].tolist() == [-84, -40, 99, 55] and a[3].tolist() == [-56, -56, 71, 71]
    assert a[-1].tolist() == [-168, -344, 183, 359]
    b = oracle.gen_base_anchors(4, [8], [0.5, 1, 2])
    assert b.tolist() == [[-21, -9, 24, 12], [-14, -14, 17, 17], [-9, -21, 12, 24]]


def test_grid_anchor_counts_and_order():
    from mxdetection_b200.synthetic import fpn_shapes
    assert sum(h * w * 3 for h, w in fpn_shapes(800, 1088)) == 217413
    assert sum(h * w * 3 for h, w in fpn_shapes(800, 1344)) == 268569
    base = oracle.gen_base_anchors(8, [8], [0.5, 1, 2])
    g = oracle.grid_anchors(base, 3, 5, 8)
    assert g.shape == (45, 4)
    assert np.array_equal(g[(2 * 5 + 4) * 3 + 1], base[1] + np.array([32, 16, 32, 16], F))
    v = oracle.valid_flags(3, 5, 2, 4, 3)
    assert v.reshape(3, 5, 3)[:, :, 0].tolist() == [[1, 1, 1, 1, 0], [1, 1, 1, 1, 0], [0, 0, 0, 0, 0]]
    ins = oracle.inside_flags(g, np.ones(45, np.uint8), 24, 40, 0)
    assert ins.sum() < 45 and np.array_equal(oracle.inside_flags(g, v, 24, 40, -1), v)


# ------------------------------------------------------------ Specs D, E ----------
def test_assign_kat4_tie_claimed_by_two_gts():
    # anchor 2 has the same IoU with GT0 and GT1 and is the best anchor of both:
    # per-anchor argmax picks the LOWER g (0), the low-quality loop lets the LATER g (1) override.
    gts = np.array([[0, 0, 9, 9], [10, 0, 19, 9], [100, 100, 149, 149]], F)
    anchors = np.array([[100, 100, 149, 149],   # == GT2 -> IoU 1 -> pos (3)
                        [300, 300, 310, 310],   # no overlap -> neg (0)
                        [5, 0, 14, 9],          # straddles GT0/GT1 equally (IoU 1/3 each)
                        [100, 100, 149, 124],   # IoU .5 with GT2: between neg and pos -> ignore (-1)
                        [0, 0, 9, 4],           # IoU .5 with GT0, not its max
                        [400, 0, 409, 9]], F)
    a, m, l = oracle.max_iou_assign(anchors, gts, np.array([7, 8, 9], np.int32), 0.7, 0.3, 0.3)
    assert m[2] == oracle.bbox_overlaps(gts[:1], anchors[2:3])[0, 0] == oracle.bbox_overlaps(gts[1:2], anchors[2:3])[0, 0]
    # GT0's best anchor is #4 (IoU .5), GT1's best is #2 (1/3): low-quality rule gives #4 -> 1, #2 -> 2
    assert a.tolist() == [3, 0, 2, -1, 1, 0]
    assert l.tolist() == [9, 0, 8, 0, 7, 0]
    # remove anchor 4: now #2 is the best of BOTH GT0 and GT1 -> later g (2) wins
    keep = [0, 1, 2, 3, 5]
    a2, _, _ = oracle.max_iou_assign(anchors[keep], gts, None, 0.7, 0.3, 0.3)
    assert a2.tolist() == [3, 0, 2, -1, 0]
    # flags: excluded anchors neither receive labels nor feed gt_max
    fl = np.array([1, 1, 0, 1, 1, 1], np.uint8)
    a3, m3, _ = oracle.max_iou_assign(anchors, gts, None, 0.7, 0.3, 0.3, flags=fl)
    assert a3[2] == -1 and m3[2] == 0 and a3[4] == 1
    # G == 0 -> everything negative
    a4, _, _ = oracle.max_iou_assign(anchors, gts[:0], None, 0.7, 0.3, 0.3)
    assert np.all(a4 == 0)
    c = cref.max_iou_assign_batch(anchors, gts[None], None, np.array([[7, 8, 9]], np.int32))
    assert c[0][0].tolist() == a.tolist() and c[2][0].tolist() == l.tolist()


def test_assign_c_vs_numpy_random():
    rng = np.random.default_rng(5)
    from mxdetection_b200.synthetic import gt_boxes
    gts = gt_boxes(rng, 400, 600, 30)
    base = oracle.gen_base_anchors(16, [8], [0.5, 1, 2])
    anchors = oracle.grid_anchors(base, 25, 38, 16)
    flags = oracle.inside_flags(anchors, np.ones(len(anchors), np.uint8), 400, 600, 0)
    a = oracle.max_iou_assign(anchors, gts, rng.integers(1, 81, 30).astype(np.int32), 0.7, 0.3, 0.3, flags=flags)
    c = cref.max_iou_assign_batch(anchors, gts[None], None, None, flags)
    assert np.array_equal(a[0], c[0][0]) and np.array_equal(a[1], c[1][0])
    assert (a[0] > 0).sum() >= 1


def test_assign_nan_propagates_like_numpy_max():
    """Spec D: 0/0 = NaN is not special-cased.  It propagates through the per-anchor and the per-GT maximum
    (numpy / torch max semantics): the anchor row stays -1, the GT takes no part in the low-quality rule."""
    gts = np.array([[50, 60, 50, 80], [10, 10, 40, 40], [100, 100, 130, 130]], F)      # GT0: zero area at delta 0
    anchors = np.array([[10, 10, 10, 10],      # zero-area anchor: 0/0 with GT0 only
                        [10, 10, 40, 40],      # == GT1
                        [100, 100, 130, 120],  # IoU 2/3 with GT2: its best
                        [300, 300, 320, 320]], F)
    a, m, _ = oracle.max_iou_assign(anchors, gts, None, 0.7, 0.3, 0.3, delta=0.0)
    assert np.isnan(m[0]) and a[0] == -1 and a[1] == 2 and a[2] == 3 and a[3] == 0
    c = cref.max_iou_assign_batch(anchors, gts[None], None, None, None, 0.7, 0.3, 0.3, 0.0)
    assert np.array_equal(c[0][0], a) and np.array_equal(c[1][0], m, equal_nan=True)
    # a NaN in GT0's row (from anchor 0) makes gt_max[0] NaN: no low-quality assignment to GT0 anywhere
    assert not (a == 1).any()


# ------------------------------------------------------------ Specs F, G ----------
def test_codec_roundtrip_and_clip():
    rng = np.random.default_rng(3)
    p = np.array([[10, 20, 110, 90], [0, 0, 15, 15], [300, 200, 420, 380]], F)
    g = p + rng.uniform(-8, 8, p.shape).astype(F)
    for stds in [(1, 1, 1, 1), (.1, .1, .2, .2)]:
        d = oracle.bbox2delta(p, g, stds=stds)
        back = oracle.delta2bbox(p, d, stds=stds)
        assert np.abs(back - g).max() < 1e-3
        assert np.array_equal(d, cref.bbox2delta(p, g, stds=stds))
        assert np.array_equal(back, cref.delta2bbox(p, d, stds=stds))
    big = np.array([[0, 0, 10, 10]], F); dl = np.array([[0, 0, 50, -50]], F)       # dw clamp at |log(16/1000)|
    out = oracle.delta2bbox(big, dl)
    assert np.isclose(out[0, 2] - out[0, 0] + 1, 11 * 1000 / 16, rtol=1e-5)
    out = oracle.delta2bbox(big, dl, max_shape=(20, 30))
    assert out[0, 0] == 0 and out[0, 2] == 29 and 0 <= out[0, 1] <= 19


def test_levels_kat3_boundaries():
    def roi(s):   # square of scale exactly s (x2-x1+1 = s)
        return np.array([[0, 0, 0, s - 1, s - 1]], F)
    assert oracle.map_roi_levels(roi(111), 4).tolist() == [0]
    assert oracle.map_roi_levels(roi(112), 4).tolist() == [1]
    assert oracle.map_roi_levels(roi(223), 4).tolist() == [1]
    assert oracle.map_roi_levels(roi(224), 4).tolist() == [2]
    assert oracle.map_roi_levels(roi(447), 4).tolist() == [2]
    assert oracle.map_roi_levels(roi(448), 4).tolist() == [3]
    assert oracle.map_roi_levels(roi(5000), 4).tolist() == [3]
    assert oracle.map_roi_levels(roi(5000), 2).tolist() == [1]
    rng = np.random.default_rng(9)
    r = np.concatenate([np.zeros((500, 1)), rng.uniform(0, 600, (500, 2)), rng.uniform(600, 1300, (500, 2))], 1).astype(F)
    assert np.array_equal(oracle.map_roi_levels(r, 4), cref.map_roi_levels(r, 4))
    # libm-free form == exact floor(log2) away from the boundaries
    w = r[:, 3] - r[:, 1] + 1; h = r[:, 4] - r[:, 2] + 1
    lit = np.clip(np.floor(np.log2(np.sqrt(w.astype(np.float64) * h) / 56 + 1e-6)), 0, 3)
    assert (oracle.map_roi_levels(r, 4) == lit).mean() > 0.99


# ---------------------------------------------------------------- Spec H ----------
def test_rpn_proposals_numpy_vs_c():
    from mxdetection_b200.synthetic import rpn_inputs
    d = rpn_inputs(2, 2, 160, 224)
    base = [oracle.gen_base_anchors(s, [8], [0.5, 1, 2]) for s in d["strides"]]
    cfg = dict(nms_pre=300, nms_thr=0.7, nms_post=100, max_num=150)
    o, n = oracle.rpn_proposals(d["scores"], d["deltas"], base, d["feat_shapes"], d["strides"], d["img_shapes"], **cfg)
    oc, nc = cref.rpn_proposals(d["scores"], d["deltas"], base, d["feat_shapes"], d["strides"], d["img_shapes"], **cfg)
    assert np.array_equal(n, nc) and np.array_equal(o, oc)
    assert n.max() == 150 and np.all(o[:, :, 4][:, :-1] >= o[:, :, 4][:, 1:])     # sorted when truncated
    assert np.all(o[..., 0] >= 0) and np.all(o[..., 2] <= 223) and np.all(o[..., 3] <= 159)
    # min-size filter drops rows
    o2, n2 = oracle.rpn_proposals(d["scores"], d["deltas"], base, d["feat_shapes"], d["strides"], d["img_shapes"],
                                  min_bbox_size=24, **cfg)
    oc2, nc2 = cref.rpn_proposals(d["scores"], d["deltas"], base, d["feat_shapes"], d["strides"], d["img_shapes"],
                                  min_bbox_size=24, **cfg)
    assert np.array_equal(o2, oc2) and np.array_equal(n2, nc2)
    w = o2[..., 2] - o2[..., 0] + 1
    assert all(np.all(w[b, :n2[b]] >= 24) for b in range(2))


def test_topk_stable_ties():
    s = np.array([.5, .9, .5, .9, .1, .5], F)
    assert oracle.topk_stable(s, 4).tolist() == [1, 3, 0, 2]
    assert oracle.topk_stable(s, 100).tolist() == [1, 3, 0, 2, 5, 4]


def test_multi_proposal_oracle_hand_cases():
    """MXNet MultiProposal restatement: anchor table (KAT-2), FilterBox growth + score -1, cyclic padding."""
    import oracle
    A = 9
    H, W = 4, 5
    cls = np.zeros((1, 2 * A, H, W), np.float32); bbox = np.zeros((1, 4 * A, H, W), np.float32)
    cls[0, A + 4, 1, 2] = 0.9          # anchor (ratio 1, scale 16) at (h=1, w=2)
    cls[0, A + 4, 1, 3] = 0.8          # its right neighbour: IoU of two 256-boxes shifted by 16 = 0.88 > 0.7
    cls[0, A + 0, 3, 0] = 0.7
    info = np.array([[H * 16.0, W * 16.0, 1.0]], np.float32)
    rois, sc = oracle.multi_proposal(cls, bbox, info, rpn_pre_nms_top_n=12, rpn_post_nms_top_n=5, threshold=0.7,
                                     rpn_min_size=16, scales=(8, 16, 32), ratios=(0.5, 1, 2), feature_stride=16)
    base = oracle.generate_anchors_mx(16, (8, 16, 32), (0.5, 1, 2))
    assert np.array_equal(base[0], [-84, -40, 99, 55]) and np.array_equal(base[4], [-120, -120, 135, 135])
    # zero deltas: proposals are the clipped anchors; the 0.8 box is suppressed by the 0.9 box
    top = rois[0]
    assert top[0] == 0 and np.array_equal(top[1:], [0, 0, W * 16 - 1, H * 16 - 1]) and sc[0, 0] == np.float32(0.9)
    assert not np.any(sc[:, 0] == np.float32(0.8))
    nk = len(np.unique(rois, axis=0))
    assert np.array_equal(rois[nk:], rois[:5 - nk]) or nk == 5          # cyclic repetition
    # a map larger than the image: positions at / beyond int(h/stride) get score -1 and sort last
    info2 = np.array([[2 * 16.0, W * 16.0, 1.0]], np.float32)
    _, sc2 = oracle.multi_proposal(cls, bbox, info2, rpn_pre_nms_top_n=-1, rpn_post_nms_top_n=3, scales=(8, 16, 32))
    assert sc2[0, 0] == np.float32(0.9) and not np.any(sc2[:, 0] == np.float32(0.7))


def test_next_row_oracles_hand_cases():
    """N1 sampling contract, N2 class-aware NMS, N4 mask target on hand-checkable inputs."""
    import oracle
    t = oracle.targets
    assigned = np.array([0, 2, -1, 1, 0, 0, 3], np.int32)
    keys = np.array([.5, .9, .99, .9, .1, .7, .2], np.float32)
    pos, neg = t.random_sample(assigned, keys, num=4, pos_fraction=0.5)
    assert pos.tolist() == [1, 3] and neg.tolist() == [5, 0]          # tie .9/.9 -> lower index first; ignore (-1) never sampled
    pos, neg = t.random_sample(assigned, keys, num=6, pos_fraction=0.5, neg_pos_ub=0)
    assert pos.tolist() == [1, 3, 6] and neg.tolist() == []
    # two identical boxes of different classes both survive; the same class is suppressed
    boxes = np.array([[0, 0, 9, 9], [0, 0, 9, 9], [50, 50, 60, 60]], np.float32)
    scores = np.array([[.1, .8, .1], [.1, .7, .6], [.9, .04, .06]], np.float32)
    dets, labels = t.multiclass_nms(boxes, scores, 0.05, 0.5, 10)
    # score order: .8 (b0,c1) keep; .7 (b1,c1) same class, IoU 1 -> out; .6 (b1,c2) keep; .1 (b0,c2) -> out; .06 (b2,c2) keep
    assert labels.tolist() == [0, 1, 1] and dets[:, 4].tolist() == [np.float32(.8), np.float32(.6), np.float32(.06)]
    # mask target of a full mask is all ones, of an empty mask all zeros
    m = np.stack([np.ones((20, 30), np.uint8), np.zeros((20, 30), np.uint8)])
    mt = t.mask_target(np.array([[2, 2, 25, 17], [2, 2, 25, 17]], np.float32), [0, 1], m, 14)
    assert mt[0].all() and not mt[1].any()


# ------------------------------------------------ size-independent properties of the oracle ----
def test_mask_paste_oracle_hand_cases():
    """Spec N4 paste: a box of exactly S x S pixels reproduces the thresholded map (sample positions hit the centres),
    a constant map stays constant at any size, boxes crossing the border are cropped, scale_factor divides the box."""
    rng = np.random.default_rng(2)
    S = 7
    m = rng.uniform(0, 1, (1, S, S)).astype(F)
    box = np.array([[10, 20, 10 + S - 1, 20 + S - 1]], F)
    out = oracle.targets.paste_masks(m, box, (40, 50))
    assert np.array_equal(out[0, 20:20 + S, 10:10 + S], (m[0] > 0.5).astype(np.uint8)) and out.sum() == (m[0] > 0.5).sum()
    ones = np.full((1, S, S), 0.75, F)
    big = oracle.targets.paste_masks(ones, np.array([[-5, -5, 30, 12]], F), (20, 25))
    assert big[0, :13, :25].all() and big[0, 13:].sum() == 0          # cropped at the border, 1 everywhere inside
    half = oracle.targets.paste_masks(ones, np.array([[20, 20, 39, 39]], F), (40, 40), scale_factor=2.0)
    assert half[0, 10:20, 10:20].all() and half.sum() == 100          # box / 2 = [10, 19]
    lab = oracle.targets.paste_masks(np.stack([np.zeros((S, S), F), ones[0]])[None], box, (40, 50), np.array([0], np.int32))
    assert lab.sum() == S * S                                          # label 0 reads channel 1


def test_oracle_properties_adjoint_linearity_idempotence():
    """The properties the GPU tests lean on at full size, checked on the oracle itself: RoIAlign backward is the
    adjoint of forward (<fwd(x), g> == <x, bwd(g)>), forward is linear in the features, NMS of its own survivors
    keeps all of them, stable top-k equals a lexsort, the sampler honours its quotas."""
    rng = np.random.default_rng(123)
    data = rng.normal(0, 1, (2, 3, 20, 28)).astype(F)
    data2 = rng.normal(0, 1, data.shape).astype(F)
    rois = np.array([[0, 3.2, 4.1, 60.7, 50.3], [1, -5, -3, 30, 40], [1, 90, 60, 111, 79], [0, 10, 10, 10, 10]], F)
    g = rng.normal(0, 1, (4, 3, 7, 7)).astype(F)
    for sr in (2, -1):
        out = oracle.roi_align_forward(data, rois, (7, 7), 0.25, sr).astype(np.float64)
        gin = oracle.roi_align_backward(g, rois, data.shape, (7, 7), 0.25, sr).astype(np.float64)
        lhs, rhs = (out * g).sum(), (data.astype(np.float64) * gin).sum()
        assert abs(lhs - rhs) <= 1e-3 * max(1.0, abs(lhs))
        both = oracle.roi_align_forward(data + data2, rois, (7, 7), 0.25, sr)
        sep = oracle.roi_align_forward(data, rois, (7, 7), 0.25, sr) + oracle.roi_align_forward(data2, rois, (7, 7), 0.25, sr)
        assert np.abs(both - sep).max() <= 1e-5
    n = 400
    xy = rng.uniform(0, 200, (n, 2)); wh = rng.uniform(4, 80, (n, 2))
    boxes = np.concatenate([xy, xy + wh], 1).astype(F)
    scores = rng.uniform(0, 1, n).astype(F)
    keep = oracle.nms(boxes, scores, 0.5, delta=1.0)
    assert np.array_equal(oracle.nms(boxes[keep], scores[keep], 0.5, delta=1.0), np.arange(len(keep)))
    s = (np.round(rng.uniform(0, 1, 5000) * 50) / 50).astype(F)
    assert np.array_equal(oracle.topk_stable(s, 300), np.lexsort((np.arange(s.size), -s.astype(np.float64)))[:300])
    assigned = rng.integers(-1, 4, 3000).astype(np.int32)
    keys = rng.random(3000).astype(F)
    for num, frac, ub in ((256, 0.5, -1), (64, 0.25, 2), (4000, 0.5, -1)):
        pos, neg = oracle.targets.random_sample(assigned, keys, num, frac, ub)
        assert len(pos) <= int(num * frac) and len(pos) + len(neg) <= num
        assert np.all(assigned[pos] > 0) and np.all(assigned[neg] == 0) and len(set(pos)) == len(pos)
        if ub >= 0:
            assert len(neg) <= ub * max(1, len(pos))
        if len(pos) < (assigned > 0).sum():            # the sample is the largest keys among the positives
            assert keys[pos].min() >= np.sort(keys[assigned > 0])[-len(pos)] - 0
