"""BASELINE config 1 split into forward / backward: python profiles/cfg1_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mxdetection_b200 import synthetic as syn
from mxdetection_b200.ops import roi_align_forward, roi_align_backward
c1 = syn.cfg1()
dev = "cuda"
rois = torch.from_numpy(c1["rois"]).to(dev)
maps = [torch.randn((1, 256, 200, 272), device=dev) for _ in range(4)]
gs = [torch.empty((1, 256, 200, 272), device=dev) for _ in range(4)]
gos = [torch.randn((512, 256, 7, 7), device=dev) for _ in range(4)]
outs = [torch.empty((512, 256, 7, 7), device=dev) for _ in range(4)]
ev = lambda: torch.cuda.Event(enable_timing=True)
for name, fn in (("fwd", lambda i: roi_align_forward(maps[i], rois, (7, 7), 0.25, 2, out=outs[i])),
                 ("bwd", lambda i: roi_align_backward(gos[i], rois, (1, 256, 200, 272), (7, 7), 0.25, 2, grad_data=gs[i]))):
    for i in range(8):
        fn(i % 4)
    torch.cuda.synchronize()
    ts = []
    for i in range(20):
        a, b = ev(), ev(); a.record(); fn(i % 4); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print(name, "ms", sum(ts) / len(ts))
