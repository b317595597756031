"""RPN proposal stage (SURVEY.md 8(a) Spec H; mmdet-0.5 RPNHead.get_proposals of
mxdetection/models/rpn_heads, /root/reference/README.md:28).

Per (image, level): top-`nms_pre` -> regenerate anchors + decode + clip ->
min-size -> NMS(`nms_thr`) -> first `nms_post`; concat levels; top-`max_num`.
The 3x3/1x1 convolutions of the head are out of scope: inputs are the
already-activated, (y,x,a)-flattened per-level scores and deltas."""
from ctypes import byref, c_int

import numpy as np
import torch

from ... import _lib as L
from ...core.anchor import AnchorGenerator


class ProposalConfig:
    """cfg of get_proposals (mmdet-0.5 test_cfg.rpn / train_cfg.rpn_proposal names)."""

    def __init__(self, nms_pre=2000, nms_post=1000, max_num=1000, nms_thr=0.7, min_bbox_size=0,
                 nms_across_levels=False):
        if nms_across_levels:
            raise NotImplementedError("nms_across_levels=True is not on the BASELINE path")
        self.nms_pre, self.nms_post, self.max_num = int(nms_pre), int(nms_post), int(max_num)
        self.nms_thr, self.min_bbox_size = float(nms_thr), float(min_bbox_size)


class RPNHead:
    def __init__(self, anchor_scales=(8,), anchor_ratios=(0.5, 1.0, 2.0), anchor_strides=(4, 8, 16, 32, 64),
                 anchor_base_sizes=None, target_means=(0.0, 0.0, 0.0, 0.0), target_stds=(1.0, 1.0, 1.0, 1.0)):
        self.anchor_strides = tuple(anchor_strides)
        self.anchor_base_sizes = tuple(anchor_strides) if anchor_base_sizes is None else tuple(anchor_base_sizes)
        self.anchor_generators = [AnchorGenerator(b, anchor_scales, anchor_ratios) for b in self.anchor_base_sizes]
        self.num_anchors = self.anchor_generators[0].num_base_anchors
        self.target_means, self.target_stds = tuple(target_means), tuple(target_stds)

    def _config(self, featmap_sizes, cfg):
        # the struct depends on the level shapes and the proposal config only: built once per shape (filling the
        # base-anchor table from Python cost 30 us per call, a quarter of the whole stage at batch 2)
        key = (tuple((int(h), int(w)) for h, w in featmap_sizes), cfg.nms_pre, cfg.nms_post, cfg.max_num,
               float(cfg.nms_thr), float(cfg.min_bbox_size), self.target_means, self.target_stds)
        cache = self.__dict__.setdefault("_cfg_cache", {})
        c = cache.get(key)
        if c is None:
            c = cache[key] = self._build_config(featmap_sizes, cfg)
        return c

    def _build_config(self, featmap_sizes, cfg):
        c = L.RpnConfig()
        c.num_levels = len(featmap_sizes)
        for l, (fh, fw) in enumerate(featmap_sizes):
            c.feat_h[l], c.feat_w[l], c.stride[l] = int(fh), int(fw), float(self.anchor_strides[l])
            base = self.anchor_generators[l].base_anchors
            for a in range(self.num_anchors):
                for j in range(4):
                    c.base_anchors[l][a][j] = float(base[a, j])
        c.num_base = self.num_anchors
        c.nms_pre, c.nms_post, c.max_num = cfg.nms_pre, cfg.nms_post, cfg.max_num
        c.nms_thr, c.min_bbox_size = cfg.nms_thr, cfg.min_bbox_size
        for j in range(4):
            c.means[j], c.stds[j] = float(self.target_means[j]), float(self.target_stds[j])
        c.delta = 1.0
        c.wh_ratio_clip = 16 / 1000
        return c

    def get_proposals(self, cls_scores, bbox_preds, featmap_sizes, img_shapes, cfg, return_workspace=False):
        """cls_scores[l] (B, H_l*W_l*A) activated; bbox_preds[l] (B, H_l*W_l*A, 4); img_shapes (B,2) int [h,w]
        (tensor on device, or a list of tuples).  Returns (proposals (B,max_num,5), num_valid (B) int32)."""
        L.require_cuda(*cls_scores, *bbox_preds)
        dev = cls_scores[0].device
        B = cls_scores[0].shape[0]
        if not isinstance(img_shapes, torch.Tensor):
            img_shapes = torch.tensor(np.asarray(img_shapes, dtype=np.int32)[:, :2], dtype=torch.int32, device=dev)
        c = self._config(featmap_sizes, cfg)
        proposals = torch.empty((B, cfg.max_num, 5), dtype=torch.float32, device=dev)
        num_valid = torch.empty((B,), dtype=torch.int32, device=dev)
        nbytes = L.lib.mxd_rpn_proposals_workspace_bytes(byref(c), B)
        if nbytes == 0 and B > 0:
            kmax, ks = c_int(), c_int()
            L.check(L.lib.mxd_rpn_proposals_dims(byref(c), byref(kmax), byref(ks)))
        ws = L.workspace(nbytes, dev, "rpn")
        s_arr, s_keep = L.dl_array([s.contiguous() for s in cls_scores])
        d_arr, d_keep = L.dl_array([d.contiguous() for d in bbox_preds])
        L.call("mxd_rpn_proposals", s_arr, d_arr, L.dl(img_shapes.contiguous()), byref(c), L.dl(proposals),
               L.dl(num_valid), ws.data_ptr(), ws.numel(), L.current_stream(dev))
        del s_keep, d_keep
        if return_workspace:
            return proposals, num_valid, (c, ws)
        return proposals, num_valid

    @staticmethod
    def stages(handle, batch):
        """Stage-wise outputs of the last get_proposals call (for parity tests):
        idx (B,L,kmax), boxes (B,L,kmax,4), keep (B,L,keep_stride), counts (B,L,2)."""
        c, ws = handle
        kmax, ks = c_int(), c_int()
        L.check(L.lib.mxd_rpn_proposals_dims(byref(c), byref(kmax), byref(ks)))
        dev = ws.device
        Lv = c.num_levels
        idx = torch.empty((batch, Lv, kmax.value), dtype=torch.int32, device=dev)
        boxes = torch.empty((batch, Lv, kmax.value, 4), dtype=torch.float32, device=dev)
        keep = torch.empty((batch, Lv, ks.value), dtype=torch.int32, device=dev)
        counts = torch.empty((batch, Lv, 2), dtype=torch.int32, device=dev)
        L.call("mxd_rpn_proposals_stages", byref(c), int(batch), ws.data_ptr(), ws.numel(), L.dl(idx), L.dl(boxes),
               L.dl(keep), L.dl(counts), L.current_stream(dev))
        return idx, boxes, keep, counts
