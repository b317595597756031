"""BASELINE config 4 mask branch (14x14) split into forward / backward: python profiles/cfg4_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mxdetection_b200 import synthetic as syn
from mxdetection_b200.ops import roi_align_fpn_forward, roi_align_fpn_backward
m = syn.cfg4_mask(batch=8, with_features=False)
dev = "cuda"
shapes = [(8, 256, h, w) for h, w in m["feat_shapes"]]
feats = [torch.randn(s, device=dev) for s in shapes]
grads = [torch.empty(s, device=dev) for s in shapes]
rois = torch.from_numpy(m["rois"]).to(dev)
go = torch.randn((rois.shape[0], 256, 14, 14), device=dev); o = torch.empty_like(go)
ev = lambda: torch.cuda.Event(enable_timing=True)
for name, fn in (("fwd", lambda: roi_align_fpn_forward(feats, rois, (14, 14), m["scales"], 2, out=o)),
                 ("bwd", lambda: roi_align_fpn_backward(go, rois, shapes, (14, 14), m["scales"], 2, grad_feats=grads))):
    for i in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for i in range(10):
        a, b = ev(), ev(); a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print(name, "ms", sum(ts) / len(ts), "rois", rois.shape[0])
