"""Oracle: anchor generation and flags (SURVEY.md 8(a) Spec C, rows C1/C2).

TEST INFRASTRUCTURE - see oracle/__init__.py.  PARITY UNPINNED by the
reference; module role = mxdetection/core/anchor (/root/reference/README.md:16).
Pinned by KAT-2 (classic 9-anchor table) and KAT-2b.
"""
import numpy as np

F = np.float32


def gen_base_anchors(base_size, scales, ratios, scale_major=True):
    """mmdet-0.5 style AnchorGenerator.gen_base_anchors -> (A,4) f32."""
    w = F(base_size); h = F(base_size)
    xc = F(0.5) * (w - F(1)); yc = F(0.5) * (h - F(1))
    ratios = np.asarray(ratios, dtype=F); scales = np.asarray(scales, dtype=F)
    hr = np.sqrt(ratios).astype(F)
    wr = (F(1) / hr).astype(F)
    if scale_major:
        ws = ((w * wr[:, None]) * scales[None, :]).astype(F).ravel()
        hs = ((h * hr[:, None]) * scales[None, :]).astype(F).ravel()
    else:
        ws = ((w * scales[:, None]) * wr[None, :]).astype(F).ravel()
        hs = ((h * scales[:, None]) * hr[None, :]).astype(F).ravel()
    base = np.stack([xc - F(0.5) * (ws - F(1)), yc - F(0.5) * (hs - F(1)),
                     xc + F(0.5) * (ws - F(1)), yc + F(0.5) * (hs - F(1))], axis=-1).astype(F)
    return np.round(base).astype(F)  # round-half-even


def generate_anchors_mx(feature_stride=16, scales=(8, 16, 32), ratios=(0.5, 1, 2)):
    """MXNet proposal.cc utils::GenerateAnchors / py-faster-rcnn table (Spec C alt)."""
    s = F(feature_stride)
    w = s; h = s
    xc = F(0.5) * (w - F(1)); yc = F(0.5) * (h - F(1))
    out = []
    for r in ratios:
        size = F(w * h)
        size_r = np.floor(F(size / F(r)))
        nw0 = np.floor(F(np.sqrt(F(size_r)) + F(0.5)))
        nh0 = np.floor(F(F(nw0 * F(r)) + F(0.5)))
        for sc in scales:
            nw = F(nw0 * F(sc)); nh = F(nh0 * F(sc))
            out.append([xc - F(0.5) * (nw - F(1)), yc - F(0.5) * (nh - F(1)),
                        xc + F(0.5) * (nw - F(1)), yc + F(0.5) * (nh - F(1))])
    return np.asarray(out, dtype=F)


def grid_anchors(base_anchors, feat_h, feat_w, stride):
    """anchor[(y*W + x)*A + a] = base[a] + [x*s, y*s, x*s, y*s] -> (H*W*A, 4)."""
    base = np.asarray(base_anchors, dtype=F)
    sx = (np.arange(feat_w, dtype=F) * F(stride)).astype(F)
    sy = (np.arange(feat_h, dtype=F) * F(stride)).astype(F)
    shift = np.zeros((feat_h, feat_w, 1, 4), dtype=F)
    shift[..., 0, 0] = sx[None, :]; shift[..., 0, 2] = sx[None, :]
    shift[..., 0, 1] = sy[:, None]; shift[..., 0, 3] = sy[:, None]
    return (shift + base[None, None]).astype(F).reshape(-1, 4)


def valid_flags(feat_h, feat_w, valid_h, valid_w, num_base):
    """mmdet AnchorGenerator.valid_flags: grid cell (y,x) valid iff y<valid_h and x<valid_w."""
    vy = np.arange(feat_h) < valid_h
    vx = np.arange(feat_w) < valid_w
    v = vy[:, None] & vx[None, :]
    return np.repeat(v.reshape(-1), num_base).astype(np.uint8)


def inside_flags(anchors, valid, img_h, img_w, allowed_border=0):
    """valid & x1>=-ab & y1>=-ab & x2<W+ab & y2<H+ab; allowed_border<0 => valid only."""
    a = np.asarray(anchors, dtype=F)
    v = np.asarray(valid).astype(bool)
    if allowed_border < 0:
        return v.astype(np.uint8)
    ab = F(allowed_border)
    ok = (a[:, 0] >= -ab) & (a[:, 1] >= -ab) & (a[:, 2] < F(img_w) + ab) & (a[:, 3] < F(img_h) + ab)
    return (v & ok).astype(np.uint8)
