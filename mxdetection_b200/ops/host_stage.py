"""Host-buffer entry of the RoI stage (forward + backward) for callers that keep tensors on the host.

In the reference the RoI stage sits between host-side NumPy/Cython target code and the MXNet engine
("GPU->host copy ... and host->GPU copy ... every iteration, per image"), so a drop-in that
is handed HOST buffers pays two PCIe crossings.  This class hides them behind the kernels instead of
serialising them: the images of a step are independent units, so image i+1 is copied in (H2D stream) while
image i is pooled and back-propagated (compute stream) and image i-1's outputs are copied out (D2H stream,
the link is full duplex).  The per-image calls are the ordinary public ops
(``roi_align_fpn_forward`` / ``roi_align_fpn_backward`` -> C ABI); nothing here is a different code path.
"""
import numpy as np
import torch

from .roi_align import roi_align_fpn_backward, roi_align_fpn_forward


class HostRoIStage:
    """Per-image software pipeline  H2D | RoIAlign fwd+bwd | D2H  over ``depth`` device slots.

    feat_shapes: per level (N, C, H, W) of the HOST feature / gradient tensors (pinned memory).
    RoIs must be grouped by image (non-decreasing batch index), as the RoI samplers emit them.
    """

    def __init__(self, feat_shapes, max_rois_per_image, pooled_size, spatial_scales, sample_ratio=2, device="cuda",
                 depth=2):
        self.dev = torch.device(device)
        self.shapes = [tuple(int(v) for v in s) for s in feat_shapes]
        self.pooled = (int(pooled_size[0]), int(pooled_size[1]))
        self.scales = [float(s) for s in spatial_scales]
        self.sr = int(sample_ratio)
        self.depth = int(depth)
        C = self.shapes[0][1]
        R = int(max_rois_per_image)
        mk = lambda *s: torch.empty(s, dtype=torch.float32, device=self.dev)   # noqa: E731
        self.slots = [dict(feats=[mk(1, C, s[2], s[3]) for s in self.shapes], grads=[mk(1, C, s[2], s[3]) for s in self.shapes],
                           rois=mk(R, 5), gout=mk(R, C, *self.pooled), out=mk(R, C, *self.pooled))
                      for _ in range(self.depth)]
        self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(self.dev) for _ in range(3))
        self.max_rois = R

    def forward_backward(self, feats_h, rois_h, grad_out_h, out_h, grads_h):
        """Enqueues the whole step and returns; ``torch.cuda.synchronize()`` (or the returned event) completes it.
        out_h (R,C,PH,PW) and grads_h[l] (N,C,H,W) are pinned host destinations."""
        b = rois_h[:, 0].numpy().astype(np.int64)
        if b.size and np.any(np.diff(b) < 0):
            raise ValueError("HostRoIStage: RoIs must be grouped by image (non-decreasing batch index)")
        N = self.shapes[0][0]
        counts = np.bincount(b, minlength=N)[:N] if b.size else np.zeros(N, np.int64)
        if counts.max(initial=0) > self.max_rois:
            raise ValueError("HostRoIStage: %d RoIs on one image, capacity %d" % (counts.max(), self.max_rois))
        starts = np.concatenate([[0], np.cumsum(counts)])
        rois0 = rois_h.clone().pin_memory()
        rois0[:, 0] = 0                                  # every slot holds one image
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.s_in, self.s_cmp, self.s_out):
            s.wait_stream(cur)
        ev_in = [None] * self.depth; ev_cmp = [None] * self.depth; ev_out = [None] * self.depth
        for i in range(N):
            k = i % self.depth
            sl = self.slots[k]
            r0, r1 = int(starts[i]), int(starts[i + 1])
            n = r1 - r0
            with torch.cuda.stream(self.s_in):
                if ev_cmp[k] is not None:
                    self.s_in.wait_event(ev_cmp[k])       # the slot's inputs were consumed
                for l, f in enumerate(feats_h):
                    sl["feats"][l].copy_(f[i:i + 1], non_blocking=True)
                if n:
                    sl["rois"][:n].copy_(rois0[r0:r1], non_blocking=True)
                    sl["gout"][:n].copy_(grad_out_h[r0:r1], non_blocking=True)
                ev_in[k] = torch.cuda.Event(); ev_in[k].record(self.s_in)
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(ev_in[k])
                if ev_out[k] is not None:
                    self.s_cmp.wait_event(ev_out[k])      # the slot's outputs were drained
                roi_align_fpn_forward(sl["feats"], sl["rois"][:n], self.pooled, self.scales, self.sr, out=sl["out"][:n])
                roi_align_fpn_backward(sl["gout"][:n], sl["rois"][:n], [g.shape for g in sl["grads"]], self.pooled,
                                       self.scales, self.sr, grad_feats=sl["grads"], accumulate=False)
                ev_cmp[k] = torch.cuda.Event(); ev_cmp[k].record(self.s_cmp)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_cmp[k])
                if n:
                    out_h[r0:r1].copy_(sl["out"][:n], non_blocking=True)
                for l, g in enumerate(grads_h):
                    g[i:i + 1].copy_(sl["grads"][l], non_blocking=True)
                ev_out[k] = torch.cuda.Event(); ev_out[k].record(self.s_out)
        for s in (self.s_in, self.s_cmp, self.s_out):
            cur.wait_stream(s)
        done = torch.cuda.Event(); done.record(cur)
        return done
