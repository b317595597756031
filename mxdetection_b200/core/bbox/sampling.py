"""Random positive / negative sampling and target packing (mmdet-0.5 RandomSampler /
anchor_target / bbox_target roles of mxdetection/core/anchor and core/bbox, /root/reference/README.md:16-17).

Device RNG contract (the reference's NumPy ``random.choice`` cannot be reproduced on a device, so the contract is
stated instead): every candidate i gets a key u_i in [0,1) (``torch.rand`` of a caller-seeded generator, or keys
passed in); the sampled positives are the positives with the LARGEST keys (ties -> lower index), likewise the
negatives.  Given the same keys the CPU oracle returns identical index lists.  Selection runs on the library's
stable top-k kernel; nothing synchronises with the host (fixed-capacity outputs, -1 padded, device counts)."""
import torch

from ... import _lib as L
from ...ops.nms import topk_stable
from .transforms import bbox2delta


class SamplingResult:
    def __init__(self, pos_inds, num_pos, neg_inds, num_neg):
        self.pos_inds, self.num_pos, self.neg_inds, self.num_neg = pos_inds, num_pos, neg_inds, num_neg


class RandomSampler:
    """RandomSampler(num, pos_fraction, neg_pos_ub=-1): at most int(num*pos_fraction) positives, negatives fill up
    to ``num`` (optionally at most neg_pos_ub * max(1, num_pos))."""

    def __init__(self, num, pos_fraction, neg_pos_ub=-1):
        self.num = int(num)
        self.pos_fraction = float(pos_fraction)
        self.neg_pos_ub = neg_pos_ub

    def sample(self, assigned_gt_inds, keys=None, generator=None):
        L.require_cuda(assigned_gt_inds, keys)
        a = assigned_gt_inds
        n = a.shape[0]
        if keys is None:
            keys = torch.rand(n, device=a.device, generator=generator)
        keys = keys.float()
        kp = min(int(self.num * self.pos_fraction), n)
        kn = min(self.num, n)
        minus = torch.full_like(keys, -1.0)
        pos_idx, pos_val = topk_stable(torch.where(a > 0, keys, minus), kp) if kp > 0 else (a.new_zeros(0), keys.new_zeros(0))
        neg_idx, neg_val = topk_stable(torch.where(a == 0, keys, minus), kn) if kn > 0 else (a.new_zeros(0), keys.new_zeros(0))
        pos_ok = pos_val >= 0
        num_pos = pos_ok.sum(dtype=torch.int32).reshape(1)
        quota = self.num - num_pos
        if self.neg_pos_ub >= 0:
            quota = torch.minimum(quota, (self.neg_pos_ub * torch.clamp(num_pos, min=1)).to(torch.int32))
        neg_ok = (neg_val >= 0) & (torch.arange(kn, device=a.device, dtype=torch.int32) < quota)
        num_neg = neg_ok.sum(dtype=torch.int32).reshape(1)
        neg1 = torch.full_like(neg_idx, -1)
        return SamplingResult(torch.where(pos_ok, pos_idx, torch.full_like(pos_idx, -1)), num_pos,
                              torch.where(neg_ok, neg_idx, neg1), num_neg)


def pack_targets(anchors, assigned_gt_inds, gt_bboxes, sampling, gt_labels=None, means=(0, 0, 0, 0), stds=(1, 1, 1, 1),
                 pos_weight=-1.0):
    """anchor_target_single / bbox_target_single packing: (labels i32 (N), label_weights (N), bbox_targets (N,4),
    bbox_weights (N,4)).  Delta encoding runs in the library's bbox2delta kernel."""
    L.require_cuda(anchors, assigned_gt_inds, gt_bboxes)
    n = anchors.shape[0]
    dev = anchors.device
    labels = torch.zeros(n, dtype=torch.int32, device=dev)
    label_w = torch.zeros(n, dtype=torch.float32, device=dev)
    tgt = torch.zeros((n, 4), dtype=torch.float32, device=dev)
    tgt_w = torch.zeros((n, 4), dtype=torch.float32, device=dev)
    pos = sampling.pos_inds.long(); neg = sampling.neg_inds.long()
    pm = pos >= 0; nm = neg >= 0
    if pos.numel():
        psafe = torch.where(pm, pos, torch.zeros_like(pos))
        g = (assigned_gt_inds.long()[psafe] - 1).clamp(min=0)
        deltas = bbox2delta(anchors[psafe].contiguous(), gt_bboxes.reshape(-1, 4)[g].contiguous(), means, stds)
        sink = torch.where(pm, pos, torch.full_like(pos, n))           # padded slots land in a discarded row
        ext = lambda t: torch.cat([t, t.new_zeros((1,) + tuple(t.shape[1:]))])   # noqa: E731
        tgt = ext(tgt).index_copy(0, sink, deltas)[:n]
        tgt_w = ext(tgt_w).index_copy(0, sink, torch.ones_like(deltas))[:n]
        lab = torch.ones_like(pos, dtype=torch.int32) if gt_labels is None else gt_labels.to(torch.int32)[g]
        labels = ext(labels).index_copy(0, sink, lab)[:n]
        w = torch.full_like(pos, 1.0 if pos_weight <= 0 else float(pos_weight), dtype=torch.float32)
        label_w = ext(label_w).index_copy(0, sink, w)[:n]
    if neg.numel():
        sink = torch.where(nm, neg, torch.full_like(neg, n))
        label_w = torch.cat([label_w, label_w.new_zeros(1)]).index_copy(0, sink, torch.ones_like(neg, dtype=torch.float32))[:n]
    return labels, label_w, tgt, tgt_w
