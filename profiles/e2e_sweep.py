"""e2e (host buffers in/out) of the RoI stage for several HostRoIStage settings: python profiles/e2e_sweep.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mxdetection_b200 import synthetic as syn
from mxdetection_b200.ops import HostRoIStage

d = syn.cfg3(batch=8, with_features=False)
shapes = [(8, 256, h, w) for h, w in d["feat_shapes"]]
feats_h = [torch.empty(s, pin_memory=True).normal_() for s in shapes]
R = d["rois"].shape[0]
gout_h = torch.empty((R, 256, 7, 7), pin_memory=True).normal_()
rois_h = torch.from_numpy(d["rois"]).pin_memory()
out_h = torch.empty((R, 256, 7, 7), pin_memory=True)
grads_h = [torch.empty(s, pin_memory=True) for s in shapes]
for split, depth in ((1, 2), (1, 3), (2, 2), (2, 3), (4, 3)):
    st = HostRoIStage(shapes, 512, (7, 7), d["scales"], 2, "cuda", depth=depth, channel_split=split)
    st.forward_backward(feats_h, rois_h, gout_h, out_h, grads_h).synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(3):
        st.forward_backward(feats_h, rois_h, gout_h, out_h, grads_h)
    t_enq = (time.perf_counter() - t0) / 3
    e1.record(); torch.cuda.synchronize()
    print("split %d depth %d: %.2f ms/step (enqueue %.2f ms)" % (split, depth, e0.elapsed_time(e1) / 3, t_enq * 1e3))
    del st
