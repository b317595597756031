"""cfg3 shard, RoIAlign forward (and backward with `bwd`) only: the short command ncu wraps.
   python profiles/fwd_only.py [fwd|bwd|both] [calls]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from mxdetection_b200 import synthetic as syn  # noqa: E402
from mxdetection_b200.ops import roi_align_fpn_backward, roi_align_fpn_forward  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "fwd"
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = "cuda"
d = syn.cfg3(batch=8, with_features=False)
shapes = [(8, 256, h, w) for h, w in d["feat_shapes"]]
feats = [torch.randn(s, device=dev) for s in shapes]
rois = torch.from_numpy(d["rois"]).to(dev)
gout = torch.randn((rois.shape[0], 256, 7, 7), device=dev)
out = torch.empty_like(gout)
grads = [torch.empty(s, device=dev) for s in shapes] if what != "fwd" else None
for _ in range(calls):
    if what in ("fwd", "both"):
        roi_align_fpn_forward(feats, rois, (7, 7), d["scales"], 2, out=out)
    if what in ("bwd", "both"):
        roi_align_fpn_backward(gout, rois, shapes, (7, 7), d["scales"], 2, grad_feats=grads)
torch.cuda.synchronize()
print("ok", float(out.flatten()[:8].sum()))
