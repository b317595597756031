"""Oracle: delta encode / decode / clip (SURVEY.md 8(a) Spec F).

TEST INFRASTRUCTURE - see oracle/__init__.py.  PARITY UNPINNED by the
reference; module role = mxdetection/core/bbox (/root/reference/README.md:17).

exp/log are CORRECTLY ROUNDED fp32 (fp64 evaluation, one rounding): libm
expf/logf results differ between glibc, numpy SIMD and CUDA by an ulp, and the
reference does not pin one; the correctly rounded value is the one every
faithful libm approximates, and it makes oracle and kernel comparable bit for
bit.  Tests still state the 1e-5 px tolerance of SURVEY.md 8(c).
"""
import numpy as np

F = np.float32


def exp_cr(x):
    return np.exp(np.asarray(x, dtype=F).astype(np.float64)).astype(F)


def log_cr(x):
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.log(np.asarray(x, dtype=F).astype(np.float64)).astype(F)


def bbox2delta(proposals, gts, means=(0, 0, 0, 0), stds=(1, 1, 1, 1)):
    p = np.asarray(proposals, dtype=F).reshape(-1, 4); g = np.asarray(gts, dtype=F).reshape(-1, 4)
    m = np.asarray(means, dtype=F); s = np.asarray(stds, dtype=F)
    px = ((p[:, 0] + p[:, 2]) * F(0.5)).astype(F); py = ((p[:, 1] + p[:, 3]) * F(0.5)).astype(F)
    pw = ((p[:, 2] - p[:, 0]) + F(1)).astype(F); ph = ((p[:, 3] - p[:, 1]) + F(1)).astype(F)
    gx = ((g[:, 0] + g[:, 2]) * F(0.5)).astype(F); gy = ((g[:, 1] + g[:, 3]) * F(0.5)).astype(F)
    gw = ((g[:, 2] - g[:, 0]) + F(1)).astype(F); gh = ((g[:, 3] - g[:, 1]) + F(1)).astype(F)
    with np.errstate(divide="ignore", invalid="ignore"):
        dx = ((gx - px) / pw).astype(F); dy = ((gy - py) / ph).astype(F)
        dw = log_cr((gw / pw).astype(F)); dh = log_cr((gh / ph).astype(F))
        d = np.stack([dx, dy, dw, dh], -1).astype(F)
        return ((d - m[None]) / s[None]).astype(F)


def delta2bbox(rois, deltas, means=(0, 0, 0, 0), stds=(1, 1, 1, 1), max_shape=None,
               wh_ratio_clip=16 / 1000):
    """rois (k,4), deltas (k,4) -> (k,4); max_shape=(h,w) clips to [0,w-1]x[0,h-1]."""
    r = np.asarray(rois, dtype=F).reshape(-1, 4); d = np.asarray(deltas, dtype=F).reshape(-1, 4)
    m = np.asarray(means, dtype=F); s = np.asarray(stds, dtype=F)
    d = ((d * s[None]).astype(F) + m[None]).astype(F)
    max_ratio = F(abs(np.log(wh_ratio_clip)))
    dx, dy = d[:, 0], d[:, 1]
    dw = np.minimum(np.maximum(d[:, 2], -max_ratio), max_ratio).astype(F)
    dh = np.minimum(np.maximum(d[:, 3], -max_ratio), max_ratio).astype(F)
    px = ((r[:, 0] + r[:, 2]) * F(0.5)).astype(F); py = ((r[:, 1] + r[:, 3]) * F(0.5)).astype(F)
    pw = ((r[:, 2] - r[:, 0]) + F(1)).astype(F); ph = ((r[:, 3] - r[:, 1]) + F(1)).astype(F)
    gw = (pw * exp_cr(dw)).astype(F); gh = (ph * exp_cr(dh)).astype(F)
    gx = (px + (pw * dx).astype(F)).astype(F); gy = (py + (ph * dy).astype(F)).astype(F)
    x1 = ((gx - (gw * F(0.5)).astype(F)) + F(0.5)).astype(F)
    y1 = ((gy - (gh * F(0.5)).astype(F)) + F(0.5)).astype(F)
    x2 = ((gx + (gw * F(0.5)).astype(F)) - F(0.5)).astype(F)
    y2 = ((gy + (gh * F(0.5)).astype(F)) - F(0.5)).astype(F)
    if max_shape is not None:
        hmax = F(max_shape[0] - 1); wmax = F(max_shape[1] - 1)
        x1 = np.minimum(np.maximum(x1, F(0)), wmax); x2 = np.minimum(np.maximum(x2, F(0)), wmax)
        y1 = np.minimum(np.maximum(y1, F(0)), hmax); y2 = np.minimum(np.maximum(y2, F(0)), hmax)
    return np.stack([x1, y1, x2, y2], -1).astype(F)
