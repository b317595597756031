"""Random positive / negative sampling and target packing (mmdet-0.5 RandomSampler /
anchor_target / bbox_target roles of mxdetection/core/anchor and core/bbox, /root/reference/README.md:16-17).

Device RNG contract (the reference's NumPy ``random.choice`` cannot be reproduced on a device, so the contract is
stated instead): every candidate i gets a key u_i in [0,1) (``torch.rand`` of a caller-seeded generator, or keys
passed in); the sampled positives are the positives with the LARGEST keys (ties -> lower index), likewise the
negatives.  Given the same keys the CPU oracle returns identical index lists.  Both calls are native
(``mxd_random_sample``: key masking + one two-segment stable top-k + a finalising CTA; ``mxd_pack_targets``: memsets +
one scatter kernel with the Spec F encoder); nothing synchronises with the host (fixed-capacity outputs, -1 padded,
device counts)."""
import torch

from ... import _lib as L


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
        a = assigned_gt_inds.to(torch.int32).contiguous()
        n = a.shape[0]
        dev = a.device
        if keys is None:
            keys = torch.rand(n, device=dev, generator=generator)
        keys = keys.float().contiguous()
        kp = min(int(self.num * self.pos_fraction), n)
        kn = min(self.num, n)
        pos = torch.empty(kp, dtype=torch.int32, device=dev)
        neg = torch.empty(kn, dtype=torch.int32, device=dev)
        counts = torch.empty(2, dtype=torch.int32, device=dev)
        ws = L.workspace(L.lib.mxd_random_sample_workspace_bytes(n, self.num), dev, "sample")
        L.call("mxd_random_sample", L.dl(a), L.dl(keys), self.num, self.pos_fraction, int(self.neg_pos_ub), L.dl(pos),
               L.dl(neg), L.dl(counts), ws.data_ptr(), ws.numel(), L.current_stream(dev))
        return SamplingResult(pos, counts[0:1], neg, counts[1:2])


def pack_targets(anchors, assigned_gt_inds, gt_bboxes, sampling, gt_labels=None, means=(0, 0, 0, 0), stds=(1, 1, 1, 1),
                 pos_weight=-1.0):
    """anchor_target_single / bbox_target_single packing: (labels i32 (N), label_weights (N), bbox_targets (N,4),
    bbox_weights (N,4)).  Delta encoding runs in the library's bbox2delta kernel."""
    L.require_cuda(anchors, assigned_gt_inds, gt_bboxes)
    n = anchors.shape[0]
    dev = anchors.device
    labels = torch.empty(n, dtype=torch.int32, device=dev)
    label_w = torch.empty(n, dtype=torch.float32, device=dev)
    tgt = torch.empty((n, 4), dtype=torch.float32, device=dev)
    tgt_w = torch.empty((n, 4), dtype=torch.float32, device=dev)
    L.call("mxd_pack_targets", L.dl(anchors.float().contiguous()), L.dl(assigned_gt_inds.to(torch.int32).contiguous()),
           L.dl(gt_bboxes.reshape(-1, 4).float().contiguous()),
           L.dl(None if gt_labels is None else gt_labels.to(torch.int32).contiguous()),
           L.dl(sampling.pos_inds.to(torch.int32).contiguous()), L.dl(sampling.neg_inds.to(torch.int32).contiguous()),
           L.float4(means), L.float4(stds), float(pos_weight), L.dl(labels), L.dl(label_w), L.dl(tgt), L.dl(tgt_w),
           L.current_stream(dev))
    return labels, label_w, tgt, tgt_w
