"""Generates tests/golden/*.npz from torchvision's compiled CPU kernels.

torchvision 0.26 CPU is an INDEPENDENT, compiled cross-oracle from the same
Caffe2/Detectron lineage as mxnet 1.3's contrib ROIAlign; it is NOT the
reference (which ships no source - see oracle/__init__.py), so parity stays
"unpinned by the reference" and these vectors pin the oracle against a second
implementation instead.  Run once, here:  python tests/golden/make_golden.py
"""
import os

import numpy as np
import torch
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))
F = np.float32


def roi_align_case(seed, N, C, H, W, R, pooled, scale, sr):
    rng = np.random.default_rng(seed)
    data = rng.standard_normal((N, C, H, W)).astype(F)
    x1 = rng.uniform(-30, W / scale, R); y1 = rng.uniform(-30, H / scale, R)
    w = np.exp(rng.uniform(np.log(4), np.log(1.2 * W / scale), R)); h = np.exp(rng.uniform(np.log(4), np.log(1.2 * H / scale), R))
    rois = np.stack([rng.integers(0, N, R), x1, y1, x1 + w, y1 + h], 1).astype(F)
    gout = rng.standard_normal((R, C) + pooled).astype(F)
    out = torch.ops.torchvision.roi_align(torch.from_numpy(data), torch.from_numpy(rois), scale, pooled[0], pooled[1], sr, False).numpy()
    gin = torch.ops.torchvision._roi_align_backward(torch.from_numpy(gout), torch.from_numpy(rois), scale, pooled[0],
                                                    pooled[1], N, C, H, W, sr, False).numpy()
    return dict(data=data, rois=rois, grad_out=gout, out=out, grad_in=gin, pooled=np.asarray(pooled),
                scale=np.asarray(scale, F), sample_ratio=np.asarray(sr))


def nms_case(seed, n, thr, span, quant):
    rng = np.random.default_rng(seed)
    xy = rng.uniform(0, span, (n, 2)); wh = rng.uniform(4, span / 3, (n, 2))
    boxes = np.concatenate([xy, xy + wh], 1).astype(F)
    scores = rng.uniform(0, 1, n)
    if quant:
        scores = np.round(scores * quant) / quant   # heavy ties
    scores = scores.astype(F)
    keep = torchvision.ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), thr).numpy().astype(np.int32)
    iou = torchvision.ops.box_iou(torch.from_numpy(boxes[:50]), torch.from_numpy(boxes[:200])).numpy()
    return dict(boxes=boxes, scores=scores, thr=np.asarray(thr, F), keep=keep, iou_50x200=iou)


if __name__ == "__main__":
    print("torchvision", torchvision.__version__)
    cases = {
        "roi_align_a": roi_align_case(11, 2, 8, 50, 68, 48, (7, 7), 0.25, 2),
        "roi_align_b": roi_align_case(12, 1, 4, 25, 34, 32, (14, 14), 0.0625, 2),
        "roi_align_c": roi_align_case(13, 2, 3, 40, 30, 24, (3, 5), 0.125, -1),
        "nms_a": nms_case(21, 600, 0.5, 200, 50),
        "nms_b": nms_case(22, 2000, 0.7, 800, 0),
        "nms_c": nms_case(23, 300, 0.3, 60, 4),
    }
    for name, c in cases.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **c)
        print(name, {k: v.shape for k, v in c.items()})
