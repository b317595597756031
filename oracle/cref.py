"""ctypes wrapper of oracle/cpu_ref.c (the plain-C restatement of Specs A-H).

TEST INFRASTRUCTURE / timed CPU baseline - see oracle/__init__.py.  Used by
tests/ (cross-check against the NumPy oracle at larger sizes) and by bench.py's
cpu_baseline / --impl reference legs (kind "port")."""
import ctypes
import os
import subprocess
from ctypes import POINTER, Structure, byref, c_double, c_float, c_int, c_ubyte, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_build", "liboracle_cpu.so")
F = np.float32


def build(force=False):
    src = os.path.join(_HERE, "cpu_ref.c")
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "_build/liboracle_cpu.so"], check=True, capture_output=True)
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(SO)
        _lib.ora_nms.restype = c_int
    return _lib


class RpnConfig(Structure):
    _fields_ = [("num_levels", c_int), ("feat_h", c_int * 8), ("feat_w", c_int * 8), ("stride", c_float * 8),
                ("num_base", c_int), ("base_anchors", ((c_float * 4) * 16) * 8),
                ("nms_pre", c_int), ("nms_post", c_int), ("max_num", c_int),
                ("nms_thr", c_float), ("min_bbox_size", c_float), ("means", c_float * 4), ("stds", c_float * 4),
                ("delta", c_float), ("wh_ratio_clip", c_double)]


def _fp(a):
    return a.ctypes.data_as(POINTER(c_float))


def _ip(a):
    return a.ctypes.data_as(POINTER(c_int))


def num_threads():
    return int(lib().ora_num_threads())


def _maps_args(maps, scales):
    maps = [np.ascontiguousarray(m, dtype=F) for m in maps]
    L = len(maps)
    ptrs = (POINTER(c_float) * L)(*[_fp(m) for m in maps])
    Hs = (c_int * L)(*[m.shape[2] for m in maps]); Ws = (c_int * L)(*[m.shape[3] for m in maps])
    sc = (c_float * L)(*[float(s) for s in scales])
    return maps, ptrs, Hs, Ws, sc


def roi_align_forward(maps, rois, pooled_size, scales, sample_ratio=2, levels=None):
    """maps: list of (N,C,H_l,W_l) (a single array = one level); scales: per-level spatial scale."""
    if isinstance(maps, np.ndarray):
        maps, scales = [maps], [scales]
    maps, ptrs, Hs, Ws, sc = _maps_args(maps, scales)
    rois = np.ascontiguousarray(rois, dtype=F)
    N, C = maps[0].shape[:2]; R = rois.shape[0]; PH, PW = pooled_size
    out = np.empty((R, C, PH, PW), dtype=F)
    lv = None if levels is None else np.ascontiguousarray(levels, dtype=np.int32)
    lib().ora_roi_align_forward(ptrs, Hs, Ws, sc, len(maps), N, C, _fp(rois), None if lv is None else _ip(lv), R,
                                PH, PW, int(sample_ratio), _fp(out))
    return out


def roi_align_backward(grad_out, rois, shapes, pooled_size, scales, sample_ratio=2, levels=None):
    single = isinstance(shapes[0], (int, np.integer))
    if single:
        shapes, scales = [shapes], [scales]
    grads = [np.zeros(tuple(s), dtype=F) for s in shapes]
    _, ptrs, Hs, Ws, sc = _maps_args(grads, scales)
    L = len(grads)
    ptrs = (POINTER(c_float) * L)(*[_fp(g) for g in grads])
    rois = np.ascontiguousarray(rois, dtype=F); grad_out = np.ascontiguousarray(grad_out, dtype=F)
    N, C = grads[0].shape[:2]; R = rois.shape[0]; PH, PW = pooled_size
    lv = None if levels is None else np.ascontiguousarray(levels, dtype=np.int32)
    lib().ora_roi_align_backward(ptrs, Hs, Ws, sc, L, N, C, _fp(rois), None if lv is None else _ip(lv), R, PH, PW,
                                 int(sample_ratio), _fp(grad_out))
    return grads[0] if single else grads


def map_roi_levels(rois, num_levels, finest_scale=56):
    rois = np.ascontiguousarray(rois, dtype=F)
    out = np.empty(rois.shape[0], np.int32)
    lib().ora_map_roi_levels(_fp(rois), rois.shape[1], rois.shape[0], int(num_levels), c_float(finest_scale), _ip(out))
    return out


def nms(boxes, scores, iou_thr, delta=0.0, topk=-1, valid_thresh=-np.inf, ids=None, force_suppress=True, max_out=-1,
        valid_mask=None):
    boxes = np.ascontiguousarray(boxes, dtype=F).reshape(-1, 4); scores = np.ascontiguousarray(scores, dtype=F)
    n = boxes.shape[0]
    keep = np.empty(max(n, 1), np.int32)
    idp = None if ids is None else _ip(np.ascontiguousarray(ids, dtype=np.int32))
    vm = None if valid_mask is None else np.ascontiguousarray(valid_mask, dtype=np.uint8)
    k = lib().ora_nms(_fp(boxes), _fp(scores), idp, n, c_float(iou_thr), c_float(delta), int(topk),
                      c_float(valid_thresh), int(bool(force_suppress)), int(max_out),
                      None if vm is None else vm.ctypes.data_as(POINTER(c_ubyte)), _ip(keep))
    return keep[:k].copy()


def bbox_overlaps(b1, b2, delta=1.0):
    b1 = np.ascontiguousarray(b1, dtype=F).reshape(-1, 4); b2 = np.ascontiguousarray(b2, dtype=F).reshape(-1, 4)
    out = np.empty((b1.shape[0], b2.shape[0]), F)
    lib().ora_bbox_overlaps(_fp(b1), b1.shape[0], _fp(b2), b2.shape[0], c_float(delta), _fp(out))
    return out


def max_iou_assign_batch(anchors, gts, num_gts=None, gt_labels=None, flags=None, pos=0.7, neg=0.3, min_pos=0.3,
                         delta=1.0):
    anchors = np.ascontiguousarray(anchors, dtype=F); gts = np.ascontiguousarray(gts, dtype=F)
    B, G = gts.shape[0], gts.shape[1]; N = anchors.shape[0]
    assigned = np.empty((B, N), np.int32); max_ov = np.empty((B, N), F); labels = np.empty((B, N), np.int32)
    ng = None if num_gts is None else np.ascontiguousarray(num_gts, dtype=np.int32)
    gl = None if gt_labels is None else np.ascontiguousarray(gt_labels, dtype=np.int32)
    fl = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint8)
    lib().ora_max_iou_assign_batch(_fp(anchors), N, _fp(gts), None if ng is None else _ip(ng), B, G,
                                   None if gl is None else _ip(gl),
                                   None if fl is None else fl.ctypes.data_as(POINTER(c_ubyte)),
                                   c_float(pos), c_float(neg), c_float(min_pos), c_float(delta),
                                   _ip(assigned), _fp(max_ov), _ip(labels))
    return assigned, max_ov, labels


def delta2bbox(rois, deltas, means=(0, 0, 0, 0), stds=(1, 1, 1, 1), max_shape=None, wh_ratio_clip=16 / 1000):
    rois = np.ascontiguousarray(rois, dtype=F); deltas = np.ascontiguousarray(deltas, dtype=F)
    out = np.empty_like(rois)
    mh, mw = (0, 0) if max_shape is None else (int(max_shape[0]), int(max_shape[1]))
    lib().ora_delta2bbox(_fp(rois), _fp(deltas), rois.shape[0], (c_float * 4)(*means), (c_float * 4)(*stds), mh, mw,
                         c_double(wh_ratio_clip), _fp(out))
    return out


def bbox2delta(p, g, means=(0, 0, 0, 0), stds=(1, 1, 1, 1)):
    p = np.ascontiguousarray(p, dtype=F); g = np.ascontiguousarray(g, dtype=F)
    out = np.empty_like(p)
    lib().ora_bbox2delta(_fp(p), _fp(g), p.shape[0], (c_float * 4)(*means), (c_float * 4)(*stds), _fp(out))
    return out


def rpn_proposals(scores, deltas, base_anchors_lvls, feat_shapes, strides, img_shapes, nms_pre=2000, nms_thr=0.7,
                  nms_post=1000, max_num=1000, min_bbox_size=0, means=(0, 0, 0, 0), stds=(1, 1, 1, 1), delta=1.0):
    L = len(scores)
    scores = [np.ascontiguousarray(s, dtype=F) for s in scores]
    deltas = [np.ascontiguousarray(d, dtype=F) for d in deltas]
    B = scores[0].shape[0]
    c = RpnConfig()
    c.num_levels = L
    for l in range(L):
        c.feat_h[l], c.feat_w[l], c.stride[l] = int(feat_shapes[l][0]), int(feat_shapes[l][1]), float(strides[l])
        base = np.asarray(base_anchors_lvls[l], dtype=F)
        for a in range(base.shape[0]):
            for j in range(4):
                c.base_anchors[l][a][j] = float(base[a, j])
    c.num_base = int(np.asarray(base_anchors_lvls[0]).shape[0])
    c.nms_pre, c.nms_post, c.max_num = int(nms_pre), int(nms_post), int(max_num)
    c.nms_thr, c.min_bbox_size = float(nms_thr), float(min_bbox_size)
    for j in range(4):
        c.means[j], c.stds[j] = float(means[j]), float(stds[j])
    c.delta, c.wh_ratio_clip = float(delta), 16 / 1000
    sp = (POINTER(c_float) * L)(*[_fp(s) for s in scores]); dp = (POINTER(c_float) * L)(*[_fp(d) for d in deltas])
    shp = np.ascontiguousarray(np.asarray(img_shapes)[:, :2], dtype=np.int32)
    out = np.empty((B, max_num, 5), F); nv = np.empty(B, np.int32)
    lib().ora_rpn_proposals(sp, dp, _ip(shp), B, byref(c), _fp(out), _ip(nv))
    return out, nv
