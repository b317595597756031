"""Oracle: RoIAlign forward / backward (SURVEY.md 8(a) Spec A).

TEST INFRASTRUCTURE - see oracle/__init__.py.  PARITY UNPINNED by the
reference (no source under /root/reference); contract = mx.nd.contrib.ROIAlign
of mxnet 1.3.0 (/root/reference/README.md:37, module mxdetection/ops,
/root/reference/README.md:24), restated per Spec A and cross-checked
bit-for-bit against torchvision's compiled CPU kernel (aligned=False) in
tests/test_oracle_roi_align.py.

Strict fp32; the accumulation order is (iy, ix) and, inside a sample,
((w1*v1 + w2*v2) + w3*v3) + w4*v4.
"""
import numpy as np

F = np.float32


def _axis_taps(start, bin_sz, grid, pooled, size):
    """Per-axis sample table shared by fwd/bwd.

    Returns (valid, lo, hi, l, h) each shaped (pooled, grid).  `valid` is the
    per-axis part of Spec A's skip test (coord < -1 or coord > size)."""
    p = np.arange(pooled, dtype=F)
    i = np.arange(grid, dtype=F)
    # y = (rsh + ph*bh) + ((iy+.5f)*bh)/gh
    c = (start + p * bin_sz)[:, None] + (((i + F(0.5)) * bin_sz) / F(grid))[None, :]
    c = c.astype(F)
    valid = ~((c < F(-1.0)) | (c > F(size)))
    c = np.where(c <= F(0), F(0), c).astype(F)
    with np.errstate(invalid="ignore"):
        lo = np.where(valid, c, F(0)).astype(np.int64)  # (int) truncation, c >= 0 here
    top = lo >= size - 1
    hi = np.where(top, size - 1, lo + 1)
    lo = np.where(top, size - 1, lo)
    c = np.where(top, lo.astype(F), c).astype(F)
    l = (c - lo.astype(F)).astype(F)
    h = (F(1.0) - l).astype(F)
    return valid, lo, hi, l, h


def _roi_geometry(roi, spatial_scale, PH, PW, sample_ratio):
    scale = F(spatial_scale)
    b = int(roi[0])
    rsw = F(roi[1]) * scale
    rsh = F(roi[2]) * scale
    rew = F(roi[3]) * scale
    reh = F(roi[4]) * scale
    rw = max(F(rew - rsw), F(1.0))
    rh = max(F(reh - rsh), F(1.0))
    bh = F(rh / F(PH))
    bw = F(rw / F(PW))
    gh = sample_ratio if sample_ratio > 0 else int(np.ceil(F(rh / F(PH))))
    gw = sample_ratio if sample_ratio > 0 else int(np.ceil(F(rw / F(PW))))
    return b, rsw, rsh, bh, bw, gh, gw


def roi_align_forward(data, rois, pooled_size, spatial_scale, sample_ratio=-1):
    """data (N,C,H,W) f32, rois (R,5) [b,x1,y1,x2,y2] -> (R,C,PH,PW) f32."""
    data = np.ascontiguousarray(data, dtype=F)
    rois = np.ascontiguousarray(rois, dtype=F)
    PH, PW = pooled_size
    N, C, H, W = data.shape
    R = rois.shape[0]
    out = np.zeros((R, C, PH, PW), dtype=F)
    for n in range(R):
        b, rsw, rsh, bh, bw, gh, gw = _roi_geometry(rois[n], spatial_scale, PH, PW, sample_ratio)
        if b < 0 or b >= N:
            continue  # Spec A: negative batch index -> zeros
        count = F(gh * gw)
        vy, yl, yh, ly, hy = _axis_taps(rsh, bh, gh, PH, H)
        vx, xl, xh, lx, hx = _axis_taps(rsw, bw, gw, PW, W)
        img = data[b]  # (C,H,W)
        acc = np.zeros((C, PH, PW), dtype=F)
        for iy in range(gh):
            for ix in range(gw):
                ok = (vy[:, iy][:, None] & vx[:, ix][None, :])  # (PH,PW)
                w1 = (hy[:, iy][:, None] * hx[:, ix][None, :]).astype(F)
                w2 = (hy[:, iy][:, None] * lx[:, ix][None, :]).astype(F)
                w3 = (ly[:, iy][:, None] * hx[:, ix][None, :]).astype(F)
                w4 = (ly[:, iy][:, None] * lx[:, ix][None, :]).astype(F)
                Y0 = yl[:, iy][:, None]; Y1 = yh[:, iy][:, None]
                X0 = xl[:, ix][None, :]; X1 = xh[:, ix][None, :]
                v1 = img[:, Y0, X0]; v2 = img[:, Y0, X1]
                v3 = img[:, Y1, X0]; v4 = img[:, Y1, X1]
                val = ((w1 * v1 + w2 * v2) + w3 * v3) + w4 * v4
                acc = np.where(ok[None], acc + val, acc).astype(F)
        out[n] = acc / count
    return out


def roi_align_backward(grad_out, rois, data_shape, pooled_size, spatial_scale,
                       sample_ratio=-1, grad_in=None):
    """grad_out (R,C,PH,PW) -> grad_data (N,C,H,W).  grad_in given => req='add'.

    Per sample g_k = (gout*w_k)/count is added at the four taps in the order
    (roi, iy, ix, tap); grad wrt rois is identically zero (not returned)."""
    grad_out = np.ascontiguousarray(grad_out, dtype=F)
    rois = np.ascontiguousarray(rois, dtype=F)
    PH, PW = pooled_size
    N, C, H, W = data_shape
    gd = np.zeros(data_shape, dtype=F) if grad_in is None else np.array(grad_in, dtype=F, copy=True)
    R = rois.shape[0]
    for n in range(R):
        b, rsw, rsh, bh, bw, gh, gw = _roi_geometry(rois[n], spatial_scale, PH, PW, sample_ratio)
        if b < 0 or b >= N:
            continue
        count = F(gh * gw)
        vy, yl, yh, ly, hy = _axis_taps(rsh, bh, gh, PH, H)
        vx, xl, xh, lx, hx = _axis_taps(rsw, bw, gw, PW, W)
        g = grad_out[n]  # (C,PH,PW)
        plane = gd[b].reshape(C, H * W)
        for iy in range(gh):
            for ix in range(gw):
                ok = (vy[:, iy][:, None] & vx[:, ix][None, :]).ravel()
                if not ok.any():
                    continue
                ws = [
                    (hy[:, iy][:, None] * hx[:, ix][None, :]),
                    (hy[:, iy][:, None] * lx[:, ix][None, :]),
                    (ly[:, iy][:, None] * hx[:, ix][None, :]),
                    (ly[:, iy][:, None] * lx[:, ix][None, :]),
                ]
                Y0 = np.broadcast_to(yl[:, iy][:, None], (PH, PW)); Y1 = np.broadcast_to(yh[:, iy][:, None], (PH, PW))
                X0 = np.broadcast_to(xl[:, ix][None, :], (PH, PW)); X1 = np.broadcast_to(xh[:, ix][None, :], (PH, PW))
                idxs = [Y0 * W + X0, Y0 * W + X1, Y1 * W + X0, Y1 * W + X1]
                for w, idx in zip(ws, idxs):
                    contrib = ((g * w.astype(F)[None]) / count).astype(F).reshape(C, -1)[:, ok]
                    np.add.at(plane, (slice(None), idx.ravel()[ok]), contrib)
    return gd


def touched_pixels(rois, data_shape, pooled_size, spatial_scale, sample_ratio=-1):
    """Boolean (N,H,W) mask of the pixels any tap of any RoI reads.

    Used by bench.py for the tighter 'unique touched px' algorithmic-bytes
    variant of SURVEY.md 8(d)."""
    rois = np.ascontiguousarray(rois, dtype=F)
    PH, PW = pooled_size
    N, C, H, W = data_shape
    m = np.zeros((N, H, W), dtype=bool)
    for n in range(rois.shape[0]):
        b, rsw, rsh, bh, bw, gh, gw = _roi_geometry(rois[n], spatial_scale, PH, PW, sample_ratio)
        if b < 0 or b >= N:
            continue
        vy, yl, yh, _, _ = _axis_taps(rsh, bh, gh, PH, H)
        vx, xl, xh, _, _ = _axis_taps(rsw, bw, gw, PW, W)
        ys = np.unique(np.concatenate([yl[vy], yh[vy]]))
        xs = np.unique(np.concatenate([xl[vx], xh[vx]]))
        m[b][np.ix_(ys, xs)] = True
    return m
