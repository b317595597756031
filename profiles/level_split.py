"""Forward (and backward) time of the cfg3 shard with the RoIs of ONE FPN level at a time: which phase of the
level-major persistent kernels is the slow one.  python profiles/level_split.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from mxdetection_b200 import synthetic as syn  # noqa: E402
from mxdetection_b200.models.roi_extractors import map_roi_levels  # noqa: E402
from mxdetection_b200.ops import roi_align_fpn_backward, roi_align_fpn_forward  # noqa: E402

dev = "cuda"
d = syn.cfg3(batch=8, with_features=False)
shapes = [(8, 256, h, w) for h, w in d["feat_shapes"]]
feats = [torch.randn(s, device=dev) for s in shapes]
grads = [torch.empty(s, device=dev) for s in shapes]
rois_all = torch.from_numpy(d["rois"]).to(dev)
lv = map_roi_levels(rois_all, 4).cpu().numpy()
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = ev(), ev()
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return round(sorted(ts)[len(ts) // 2], 4)


res = {}
for name, sel in [("all", np.ones(len(lv), bool)), ("none", np.zeros(len(lv), bool)), ("L123", lv > 0), ("L01", lv < 2), ("L23", lv >= 2)] + [("L%d" % l, lv == l) for l in range(4)]:
    rois = rois_all[torch.from_numpy(np.nonzero(sel)[0]).to(dev)].contiguous()
    gout = torch.randn((rois.shape[0], 256, 7, 7), device=dev)
    out = torch.empty_like(gout)
    f = timed(lambda: roi_align_fpn_forward(feats, rois, (7, 7), d["scales"], 2, out=out))
    b = timed(lambda: roi_align_fpn_backward(gout, rois, shapes, (7, 7), d["scales"], 2, grad_feats=grads, accumulate=True))
    bw = timed(lambda: roi_align_fpn_backward(gout, rois, shapes, (7, 7), d["scales"], 2, grad_feats=grads, accumulate=False))
    res[name] = {"rois": int(rois.shape[0]), "fwd_ms": f, "bwd_add_ms": b, "bwd_write_ms": bw}
print(json.dumps(res))
