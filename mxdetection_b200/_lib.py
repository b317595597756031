"""ctypes binding of libmxdet_sm100.so and DLPack tensor exchange.

Tensors cross the C ABI as borrowed ``DLTensor*`` taken from the producer's
DLPack capsule (``obj.__dlpack__()`` / ``torch.utils.dlpack.to_dlpack``); the
capsule is kept alive for the duration of the (asynchronous-enqueue) call and
is never consumed, so its own destructor releases it.

There is NO CPU fallback: if the CUDA library is missing, importing this module
raises; if a CPU tensor is passed, the library answers MXD_ENOTSUP and
``MXDetError`` is raised.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, byref, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t,
                    c_uint8, c_uint16, c_uint64, c_void_p)

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmxdet_sm100.so")

MXD_MAX_LEVELS = 8
MXD_MAX_BASE_ANCHORS = 16
MXD_SORT_CAP = 8192


class MXDetError(RuntimeError):
    """Raised when a libmxdet_sm100 entry point returns a negative MXD_E* code
    (the counterpart of MXNetError raised from MXGetLastError())."""

    def __init__(self, code, msg):
        super().__init__("libmxdet_sm100 error %d: %s" % (code, msg))
        self.code = code


class DLDevice(Structure):
    _fields_ = [("device_type", c_int), ("device_id", c_int32)]


class DLDataType(Structure):
    _fields_ = [("code", c_uint8), ("bits", c_uint8), ("lanes", c_uint16)]


class DLTensor(Structure):
    _fields_ = [("data", c_void_p), ("device", DLDevice), ("ndim", c_int32), ("dtype", DLDataType),
                ("shape", POINTER(c_int64)), ("strides", POINTER(c_int64)), ("byte_offset", c_uint64)]


class RpnConfig(Structure):
    """mxd_rpn_config of include/mxdet.h."""
    _fields_ = [("num_levels", c_int),
                ("feat_h", c_int * MXD_MAX_LEVELS),
                ("feat_w", c_int * MXD_MAX_LEVELS),
                ("stride", c_float * MXD_MAX_LEVELS),
                ("num_base", c_int),
                ("base_anchors", ((c_float * 4) * MXD_MAX_BASE_ANCHORS) * MXD_MAX_LEVELS),
                ("nms_pre", c_int), ("nms_post", c_int), ("max_num", c_int),
                ("nms_thr", c_float), ("min_bbox_size", c_float),
                ("means", c_float * 4), ("stds", c_float * 4),
                ("delta", c_float), ("wh_ratio_clip", c_double)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libmxdet_sm100.so is not built (expected at %s). Run `python mxdetection_b200/build.py` "
            "(needs nvcc with sm_100a support). There is no CPU fallback." % LIB_PATH)
    return ctypes.CDLL(LIB_PATH)


lib = _load()
lib.mxd_last_error.restype = c_char_p
lib.mxd_launch_count.restype = c_uint64
for _n in ("mxd_topk_stable_workspace_bytes", "mxd_nms_workspace_bytes", "mxd_nms_batched_workspace_bytes", "mxd_box_nms_workspace_bytes",
           "mxd_max_iou_assign_workspace_bytes", "mxd_rpn_proposals_workspace_bytes",
           "mxd_roi_align_workspace_bytes", "mxd_multi_proposal_workspace_bytes", "mxd_random_sample_workspace_bytes", "mxd_det_bboxes_workspace_bytes"):
    if hasattr(lib, _n):
        getattr(lib, _n).restype = c_size_t

_PyCapsule_GetPointer = ctypes.pythonapi.PyCapsule_GetPointer
_PyCapsule_GetPointer.restype = c_void_p
_PyCapsule_GetPointer.argtypes = [ctypes.py_object, c_char_p]


class Borrowed:
    """A DLTensor* borrowed from a DLPack capsule; keeps the capsule alive."""
    __slots__ = ("capsule", "ptr")

    def __init__(self, obj):
        if isinstance(obj, torch.Tensor):
            self.capsule = torch.utils.dlpack.to_dlpack(obj)
        else:
            self.capsule = obj.__dlpack__()
        # DLManagedTensor starts with its DLTensor, so the capsule pointer is the DLTensor*.
        self.ptr = c_void_p(_PyCapsule_GetPointer(self.capsule, b"dltensor"))


def dl(obj):
    """DLTensor* argument (None -> NULL)."""
    return None if obj is None else Borrowed(obj)


def dl_array(objs):
    """`const DLTensor* const*` table from a sequence of tensors."""
    borrowed = [Borrowed(o) for o in objs]
    arr = (c_void_p * len(borrowed))(*[b.ptr for b in borrowed])
    return arr, borrowed


def check(rc):
    if rc != 0:
        raise MXDetError(rc, lib.mxd_last_error().decode("utf-8", "replace"))


def call(name, *args):
    """Invoke an entry point; Borrowed args are unwrapped and kept alive until it returns."""
    raw = [a.ptr if isinstance(a, Borrowed) else a for a in args]
    check(getattr(lib, name)(*raw))


def current_stream(device=None):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def launch_count():
    return int(lib.mxd_launch_count())


def float4(vals):
    return (c_float * 4)(*[float(v) for v in vals])


_WORKSPACES = {}


def workspace(nbytes, device, tag):
    """A cached, growing, 256-byte aligned device scratch buffer per (device, stream, tag)."""
    key = (str(device), torch.cuda.current_stream(device).cuda_stream, tag)
    buf = _WORKSPACES.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _WORKSPACES[key] = buf
    return buf


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and isinstance(t, torch.Tensor) and not t.is_cuda:
            raise MXDetError(-2, "tensor on %s: libmxdet_sm100 has no CPU fallback" % t.device)


# ---- argtypes (explicit so that Python floats become C floats, not doubles) ----
_P = c_void_p
_SIG = {
    "mxd_roi_align_workspace_bytes": [c_int, c_int, c_int, c_int, POINTER(c_int), POINTER(c_int), c_int, c_int, c_int],
    "mxd_roi_align_forward": [_P, _P, _P, c_int, c_int, c_float, c_int, _P, c_size_t, _P],
    "mxd_roi_align_backward": [_P, _P, _P, c_int, c_int, c_float, c_int, c_int, _P, c_size_t, _P],
    "mxd_map_roi_levels": [_P, _P, c_int, c_float, _P],
    "mxd_roi_align_fpn_forward": [_P, c_int, POINTER(c_float), _P, _P, _P, c_int, c_int, c_int, c_float, _P, c_size_t, _P],
    "mxd_roi_align_fpn_backward": [_P, _P, _P, _P, c_int, POINTER(c_float), c_int, c_int, c_int, c_float, c_int, _P,
                                   c_size_t, _P],
    "mxd_topk_stable_workspace_bytes": [c_int, c_int, c_int],
    "mxd_topk_stable": [_P, _P, _P, c_int, _P, c_size_t, _P],
    "mxd_nms_workspace_bytes": [c_int, c_int],
    "mxd_nms": [_P, _P, _P, _P, _P, c_float, c_float, c_int, c_float, c_int, c_int, _P, c_size_t, _P],
    "mxd_nms_batched_workspace_bytes": [c_int, c_int, c_int],
    "mxd_nms_batched": [_P, _P, _P, _P, c_int, _P, _P, c_float, c_float, c_int, c_float, c_int, c_int, _P, c_size_t, _P],
    "mxd_box_nms_workspace_bytes": [c_int, c_int, c_int],
    "mxd_box_nms": [_P, _P, _P, c_float, c_float, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_size_t, _P],
    "mxd_box_nms_backward": [_P, _P, _P, _P],
    "mxd_grid_anchors": [POINTER(c_float), c_int, c_int, c_int, c_float, _P, _P],
    "mxd_valid_flags": [c_int, c_int, c_int, c_int, c_int, _P, _P],
    "mxd_inside_flags": [_P, _P, c_int, c_int, c_float, _P, _P],
    "mxd_bbox_overlaps": [_P, _P, _P, c_float, _P],
    "mxd_max_iou_assign_workspace_bytes": [c_int, c_int],
    "mxd_max_iou_assign": [_P, _P, _P, _P, _P, _P, _P, _P, c_float, c_float, c_float, c_float, _P, c_size_t, _P],
    "mxd_bbox2delta": [_P, _P, _P, POINTER(c_float), POINTER(c_float), _P],
    "mxd_det_bboxes_workspace_bytes": [ctypes.c_longlong, c_int, c_int],
    "mxd_det_bboxes": [_P, _P, _P, POINTER(c_float), POINTER(c_float), c_int, c_int, c_double, c_float, c_float, c_float, c_float,
                       c_int, _P, _P, _P, _P, c_size_t, _P],
    "mxd_random_sample_workspace_bytes": [ctypes.c_longlong, c_int],
    "mxd_random_sample": [_P, _P, c_int, c_double, c_int, _P, _P, _P, _P, c_size_t, _P],
    "mxd_pack_targets": [_P, _P, _P, _P, _P, _P, POINTER(c_float), POINTER(c_float), c_float, _P, _P, _P, _P, _P],
    "mxd_delta2bbox": [_P, _P, _P, POINTER(c_float), POINTER(c_float), c_int, c_int, c_double, _P],
    "mxd_rpn_proposals_workspace_bytes": [POINTER(RpnConfig), c_int],
    "mxd_rpn_proposals_dims": [POINTER(RpnConfig), POINTER(c_int), POINTER(c_int)],
    "mxd_rpn_proposals": [_P, _P, _P, POINTER(RpnConfig), _P, _P, _P, c_size_t, _P],
    "mxd_rpn_proposals_stages": [POINTER(RpnConfig), c_int, _P, c_size_t, _P, _P, _P, _P, _P],
    "mxd_mask_target": [_P, _P, _P, _P, c_int, c_int, c_float, _P],
    "mxd_paste_masks": [_P, _P, _P, _P, c_float, c_float, _P],
    "mxd_pack_detections": [_P, _P, c_int, _P, _P],
    "mxd_copy2d_async": [_P, c_size_t, _P, c_size_t, c_size_t, c_size_t, c_int, _P],
    "mxd_multi_proposal_workspace_bytes": [c_int, c_int, c_int, c_int, c_int, c_int],
    "mxd_multi_proposal": [_P, _P, _P, _P, _P, POINTER(c_float), c_int, c_float, c_int, c_int, c_float, c_float, _P,
                           c_size_t, _P],
}
for _n, _a in _SIG.items():
    getattr(lib, _n).argtypes = _a
