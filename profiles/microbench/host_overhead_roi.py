"""Host-side cost of one roi_align_fpn_forward / backward call through the Python mirror (ctypes + DLPack), measured
by enqueueing calls on a tiny workload (the GPU work is negligible, the host path identical to the benchmark's).
   python profiles/microbench/host_overhead_roi.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
from mxdetection_b200 import synthetic as syn  # noqa: E402
from mxdetection_b200.ops import roi_align_fpn_backward, roi_align_fpn_forward  # noqa: E402

dev = "cuda"
d = syn.cfg3(batch=8, with_features=False)
shapes = [(8, 256, h, w) for h, w in d["feat_shapes"]]
feats = [torch.zeros(s, device=dev) for s in shapes]
grads = [torch.empty(s, device=dev) for s in shapes]
for R in (8, 4096):
    rois = torch.from_numpy(d["rois"][:R]).to(dev)
    gout = torch.zeros((R, 256, 7, 7), device=dev)
    out = torch.empty_like(gout)
    calls = {"forward": lambda: roi_align_fpn_forward(feats, rois, (7, 7), d["scales"], 2, out=out),
             "backward": lambda: roi_align_fpn_backward(gout, rois, shapes, (7, 7), d["scales"], 2, grad_feats=grads)}
    for name, fn in calls.items():
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        n = 100
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print("R=%d %s: host enqueue %.1f us/call, wall incl. GPU %.1f us/call" % (R, name, (t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6))
