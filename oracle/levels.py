"""Oracle: FPN level assignment (SURVEY.md 8(a) Spec G).

TEST INFRASTRUCTURE - see oracle/__init__.py.  PARITY UNPINNED by the
reference; module role = mxdetection/models/roi_extractors
(/root/reference/README.md:32).  Canonical libm-free threshold form:
floor(log2(v)) clamped to [0,L-1] == [v>=2]+[v>=4]+[v>=8]+... (exact floor).
"""
import numpy as np

F = np.float32


def map_roi_levels(rois, num_levels, finest_scale=56):
    r = np.asarray(rois, dtype=F)
    b = r[:, 1:5] if r.shape[1] == 5 else r
    w = ((b[:, 2] - b[:, 0]) + F(1)).astype(F); h = ((b[:, 3] - b[:, 1]) + F(1)).astype(F)
    with np.errstate(invalid="ignore"):
        s = np.sqrt((w * h).astype(F)).astype(F)
    v = ((s / F(finest_scale)).astype(F) + F(1e-6)).astype(F)
    lvl = np.zeros(len(b), np.int32)
    t = F(2)
    for _ in range(num_levels - 1):
        lvl += (v >= t).astype(np.int32)   # NaN compares false -> level 0
        t = F(t * F(2))
    return lvl
