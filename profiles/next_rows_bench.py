"""Timings of the SURVEY 8(f) "next" rows at the sizes the reference pipeline runs them:
python profiles/next_rows_bench.py  (events around back-to-back calls, synthetic inputs, one GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mxdetection_b200 import synthetic as syn
from mxdetection_b200.core.bbox import MaxIoUAssigner, RandomSampler, pack_targets
from mxdetection_b200.core.mask import mask_target
from mxdetection_b200.models.bbox_heads import get_det_bboxes
from mxdetection_b200.models.rpn_heads import MultiProposal

F = np.float32
dev = "cuda"
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
rng = np.random.default_rng(3)


def timed(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


# N1: sample 256 of the 268 569 anchors of one 800x1344 image (after assignment) and pack label/target tensors
d = syn.assigner_inputs(4, 1)
from mxdetection_b200.core.anchor import AnchorGenerator
anchors = torch.cat([AnchorGenerator(s, [8], [0.5, 1.0, 2.0]).grid_anchors((fh, fw), s) for (fh, fw), s in zip(d["feat_shapes"], d["strides"])])
gts = T(d["gts"][0][: int(d["num_gts"][0])])
asg = MaxIoUAssigner(0.7, 0.3, 0.3).assign(anchors, gts)
keys = torch.rand(anchors.shape[0], device=dev)
sampler = RandomSampler(256, 0.5, -1)


def n1():
    s = sampler.sample(asg.gt_inds, keys)
    pack_targets(anchors, asg.gt_inds, gts, s)


print("N1 sample 256 of %d anchors + pack targets: %.3f ms" % (anchors.shape[0], timed(n1)))

# N2: detection post-processing, 1000 proposals x 81 classes -> top 100
n, C = 1000, 81
rois = np.concatenate([np.zeros((n, 1)), syn.gt_boxes(rng, 800, 1344, n)], 1).astype(F)
logits = rng.normal(0, 2, (n, C)); logits[:, 0] += 3
score = (np.exp(logits) / np.exp(logits).sum(1, keepdims=True)).astype(F)
pred = rng.normal(0, 1.0, (n, 4 * C)).astype(F)
tr, ts, tp = T(rois), T(score), T(pred)
print("N2 get_det_bboxes 1000 x 81 classes -> 100: %.3f ms" % timed(lambda: get_det_bboxes(tr, ts, tp, (800, 1344), 1.0, 0.05, 0.5, 100)))

# N3: mx.nd.contrib.MultiProposal, 2 images, 50x84 stride-16 map, 12 anchors/cell, 6000 -> 300
N_, H, W, A = 2, 50, 84, 12
cls = (1.0 / (1.0 + np.exp(-rng.normal(-2, 2, (N_, 2 * A, H, W))))).astype(F)
bbox = rng.normal(0, 0.4, (N_, 4 * A, H, W)).astype(F)
info = np.array([[800.0, 1344.0, 1.0]] * N_, F)
tc, tb, ti = T(cls), T(bbox), T(info)
kw = dict(rpn_pre_nms_top_n=6000, rpn_post_nms_top_n=300, threshold=0.7, rpn_min_size=16, scales=(4, 8, 16, 32),
          ratios=(0.5, 1, 2), feature_stride=16)
print("N3 MultiProposal 2 x %d anchors, 6000 -> 300: %.3f ms" % (H * W * A, timed(lambda: MultiProposal(tc, tb, ti, **kw))))

# N4: mask targets, 128 positive RoIs, 20 GT masks of 800x1344, 28x28
G = 20
masks = (rng.random((G, 800, 1344)) > 0.5).astype(np.uint8)
props = syn.gt_boxes(rng, 800, 1344, 128)
inds = rng.integers(0, G, 128).astype(np.int32)
tm, tpz, tix = T(masks), T(props), T(inds)
print("N4 mask_target 128 RoIs x 28x28 from 20 masks: %.3f ms" % timed(lambda: mask_target(tpz, tix, tm, 28)))
