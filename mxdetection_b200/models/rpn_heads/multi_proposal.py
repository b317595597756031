"""``mx.nd.contrib.MultiProposal`` / ``Proposal`` of mxnet 1.3.0 - the single-level NCHW proposal operator that an
MXNet RPN head (mxdetection/models/rpn_heads, /root/reference/README.md:28) calls.  Semantics: SURVEY.md 8(a)
Spec H alt-mode / Spec F MX13 variant (multi_proposal.cc): +1 box convention, no dw/dh clamp, min-size boxes get
score -1, stable score order, strict ``iou > threshold``, output cyclically padded to ``rpn_post_nms_top_n`` rows.
"""
import ctypes

import numpy as np
import torch

from ... import _lib as L
from ...core.anchor.anchor_generator import generate_anchors_mx


def MultiProposal(cls_prob, bbox_pred, im_info, rpn_pre_nms_top_n=6000, rpn_post_nms_top_n=300, threshold=0.7,
                  rpn_min_size=16, scales=(4, 8, 16, 32), ratios=(0.5, 1, 2), feature_stride=16, output_score=False):
    """cls_prob (N,2A,H,W), bbox_pred (N,4A,H,W), im_info (N,3) [h,w,scale] -> rois (N*post_n,5) [+ scores (N*post_n,1)]."""
    L.require_cuda(cls_prob, bbox_pred, im_info)
    cls_prob = cls_prob.contiguous(); bbox_pred = bbox_pred.contiguous(); im_info = im_info.contiguous()
    base = np.ascontiguousarray(generate_anchors_mx(feature_stride, scales, ratios), dtype=np.float32)
    A = base.shape[0]
    N, _, H, W = cls_prob.shape
    post_n = int(rpn_post_nms_top_n)
    rois = torch.empty((N * post_n, 5), dtype=torch.float32, device=cls_prob.device)
    scores = torch.empty((N * post_n, 1), dtype=torch.float32, device=cls_prob.device)
    nbytes = L.lib.mxd_multi_proposal_workspace_bytes(int(N), int(A), int(H), int(W), int(rpn_pre_nms_top_n), post_n)
    ws = L.workspace(nbytes, cls_prob.device, "multi_proposal")
    flat = base.reshape(-1)
    L.call("mxd_multi_proposal", L.dl(cls_prob), L.dl(bbox_pred), L.dl(im_info), L.dl(rois), L.dl(scores),
           (ctypes.c_float * flat.size)(*flat.tolist()), int(A), float(feature_stride), int(rpn_pre_nms_top_n), post_n,
           float(threshold), float(rpn_min_size), ws.data_ptr(), ws.numel(), L.current_stream(cls_prob.device))
    return (rois, scores) if output_score else rois


def Proposal(cls_prob, bbox_pred, im_info, **kw):
    """``mx.nd.contrib.Proposal``: the batch-1 form of MultiProposal (same arithmetic)."""
    return MultiProposal(cls_prob, bbox_pred, im_info, **kw)
