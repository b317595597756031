"""Quick split timing of the RoIAlign kernels (CUDA events, inputs larger than L2 or rotated):
   python profiles/roi_bench.py [iters]
BASELINE config 3 shard (8 x 512 RoIs, 4 FPN maps, 7x7), config 1 (one 256x200x272 map, 4 rotating buffer sets),
config 4a (14x14 mask branch, 1024 RoIs).  Prints one JSON line; bench.py remains the judged number."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import mxdetection_b200 as m  # noqa: E402
from mxdetection_b200 import synthetic as syn  # noqa: E402
from mxdetection_b200.ops import roi_align_backward, roi_align_forward, roi_align_fpn_backward, roi_align_fpn_forward  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = "cuda"
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731


def timed(fn, n=iters, warm=4):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    ts = []
    for i in range(n):
        a, b = ev(), ev()
        a.record(); fn(i); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return sum(ts) / len(ts), ts[len(ts) // 2], ts[0]


res = {}
d = syn.cfg3(batch=8, with_features=False)
shapes = [(8, 256, h, w) for h, w in d["feat_shapes"]]
feats = [torch.randn(s, device=dev) for s in shapes]
rois = torch.from_numpy(d["rois"]).to(dev)
gout = torch.randn((rois.shape[0], 256, 7, 7), device=dev)
out = torch.empty_like(gout)
grads = [torch.empty(s, device=dev) for s in shapes]
l0 = m.launch_count()
roi_align_fpn_forward(feats, rois, (7, 7), d["scales"], 2, out=out)
res["fwd_launches"] = m.launch_count() - l0
res["cfg3_fwd_ms"] = timed(lambda i: roi_align_fpn_forward(feats, rois, (7, 7), d["scales"], 2, out=out))
res["cfg3_bwd_ms"] = timed(lambda i: roi_align_fpn_backward(gout, rois, shapes, (7, 7), d["scales"], 2, grad_feats=grads))
dm = syn.cfg4_mask(batch=8, with_features=False)
rois_m = torch.from_numpy(dm["rois"]).to(dev)
go_m = torch.randn((rois_m.shape[0], 256, 14, 14), device=dev)
o_m = torch.empty_like(go_m)
res["cfg4a_fwd_ms"] = timed(lambda i: roi_align_fpn_forward(feats, rois_m, (14, 14), d["scales"], 2, out=o_m))
res["cfg4a_bwd_ms"] = timed(lambda i: roi_align_fpn_backward(go_m, rois_m, shapes, (14, 14), d["scales"], 2, grad_feats=grads))
del feats, grads, gout, out, go_m, o_m
c1 = syn.cfg1()
rois1 = torch.from_numpy(c1["rois"]).to(dev)
maps = [torch.randn((1, 256, 200, 272), device=dev) for _ in range(4)]
gs = [torch.empty((1, 256, 200, 272), device=dev) for _ in range(4)]
gos = [torch.randn((512, 256, 7, 7), device=dev) for _ in range(4)]
outs = [torch.empty((512, 256, 7, 7), device=dev) for _ in range(4)]
res["cfg1_fwd_ms"] = timed(lambda i: roi_align_forward(maps[i % 4], rois1, (7, 7), 0.25, 2, out=outs[i % 4]))
res["cfg1_bwd_ms"] = timed(lambda i: roi_align_backward(gos[i % 4], rois1, (1, 256, 200, 272), (7, 7), 0.25, 2, grad_data=gs[i % 4]))
print(json.dumps({k: ([round(x, 4) for x in v] if isinstance(v, tuple) else v) for k, v in res.items()}))
