"""Anchor generation (SURVEY.md 8(a) Spec C; mmdet-0.5 AnchorGenerator role of
mxdetection/core/anchor, /root/reference/README.md:16).

The A base anchors are a tiny host-side table (strict fp32, round-half-even);
the H*W*A grid is produced on the device (or regenerated inside the proposal
kernel and never stored).
"""
import ctypes

import numpy as np
import torch

from ... import _lib as L

_F = np.float32


def generate_anchors_mx(feature_stride=16, scales=(8, 16, 32), ratios=(0.5, 1, 2)):
    """MXNet proposal.cc GenerateAnchors (py-faster-rcnn table), ratio-major then scale -> (A,4) f32."""
    s = _F(feature_stride)
    ctr = _F(0.5) * (s - _F(1))
    rows = []
    for r in ratios:
        size_r = np.floor(_F(_F(s * s) / _F(r)))
        w0 = np.floor(_F(np.sqrt(_F(size_r)) + _F(0.5)))
        h0 = np.floor(_F(_F(w0 * _F(r)) + _F(0.5)))
        for sc in scales:
            w = _F(w0 * _F(sc)); h = _F(h0 * _F(sc))
            hw = _F(0.5) * (w - _F(1)); hh = _F(0.5) * (h - _F(1))
            rows.append([ctr - hw, ctr - hh, ctr + hw, ctr + hh])
    return np.asarray(rows, dtype=_F)


class AnchorGenerator:
    """AnchorGenerator(base_size, scales, ratios, scale_major=True) as in mmdet 0.5."""

    def __init__(self, base_size, scales, ratios, scale_major=True):
        self.base_size = base_size
        self.scales = np.asarray(scales, dtype=_F)
        self.ratios = np.asarray(ratios, dtype=_F)
        self.scale_major = scale_major
        self.base_anchors = self.gen_base_anchors()

    @property
    def num_base_anchors(self):
        return self.base_anchors.shape[0]

    def gen_base_anchors(self):
        side = _F(self.base_size)
        ctr = _F(0.5) * (side - _F(1))
        h_r = np.sqrt(self.ratios).astype(_F)
        w_r = (_F(1) / h_r).astype(_F)
        if self.scale_major:
            ws = ((side * w_r[:, None]) * self.scales[None, :]).astype(_F).reshape(-1)
            hs = ((side * h_r[:, None]) * self.scales[None, :]).astype(_F).reshape(-1)
        else:
            ws = ((side * self.scales[:, None]) * w_r[None, :]).astype(_F).reshape(-1)
            hs = ((side * self.scales[:, None]) * h_r[None, :]).astype(_F).reshape(-1)
        half_w = _F(0.5) * (ws - _F(1)); half_h = _F(0.5) * (hs - _F(1))
        table = np.stack([ctr - half_w, ctr - half_h, ctr + half_w, ctr + half_h], axis=-1).astype(_F)
        return np.round(table).astype(_F)

    def _base_ptr(self):
        flat = np.ascontiguousarray(self.base_anchors, dtype=_F).reshape(-1)
        return (ctypes.c_float * flat.size)(*flat.tolist())

    def grid_anchors(self, featmap_size, stride=16, device="cuda"):
        """(feat_h, feat_w) -> (feat_h*feat_w*A, 4) f32 on `device`, order (y,x,a)."""
        fh, fw = featmap_size
        out = torch.empty((fh * fw * self.num_base_anchors, 4), dtype=torch.float32, device=device)
        L.call("mxd_grid_anchors", self._base_ptr(), self.num_base_anchors, int(fh), int(fw), float(stride),
               L.dl(out), L.current_stream(out.device))
        return out

    def valid_flags(self, featmap_size, valid_size, device="cuda"):
        fh, fw = featmap_size
        vh, vw = valid_size
        out = torch.empty((fh * fw * self.num_base_anchors,), dtype=torch.uint8, device=device)
        L.call("mxd_valid_flags", int(fh), int(fw), int(vh), int(vw), self.num_base_anchors, L.dl(out),
               L.current_stream(out.device))
        return out
