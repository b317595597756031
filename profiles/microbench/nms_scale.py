"""NMS kernels against the segment length (one segment): run under
ncu --metrics gpu__time_duration.sum to read nms_mask / nms_scan per n.  python profiles/microbench/nms_scale.py"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from mxdetection_b200.ops import nms_indices

for n in (512, 1024, 2048, 4096, 8192):
    rng = np.random.default_rng(n)
    span = 40 * math.sqrt(n) + 20
    xy = rng.uniform(0, span, (n, 2)); wh = rng.uniform(4, 120, (n, 2))
    boxes = torch.from_numpy(np.concatenate([xy, xy + wh], 1).astype(np.float32)).cuda()
    scores = torch.from_numpy(rng.uniform(0, 1, n).astype(np.float32)).cuda()
    for _ in range(3):
        keep, num = nms_indices(boxes, scores, 0.7, delta=1.0)
    torch.cuda.synchronize()
    print(n, int(num.item()))
