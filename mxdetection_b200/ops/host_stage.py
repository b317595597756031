"""Host-buffer entry of the RoI stage (forward + backward) for callers that keep tensors on the host.

In the reference the RoI stage sits between host-side NumPy/Cython target code and the MXNet engine
("GPU->host copy ... and host->GPU copy ... every iteration, per image"), so a drop-in that
is handed HOST buffers pays two PCIe crossings.  This class hides them behind the kernels instead of
serialising them: (image, channel slice) units are independent, so unit u+1 is copied in (H2D stream) while
unit u is pooled and back-propagated (compute stream) and unit u-1's outputs are copied out (D2H stream, the
link is full duplex).  ``channel_split`` > 1 cuts every image into channel slices to shorten the two exposed
ends of the pipeline (the first H2D and the last D2H); on the B200 boxes measured (profiles/e2e_sweep.py) the
smaller, strided transfers cost more PCIe efficiency than that saves (22.4 ms whole images, 23.4 ms halves,
25.5 ms quarters for BASELINE config 3), so the default is whole images.  The per-unit calls are the ordinary public ops
(``roi_align_fpn_forward`` / ``roi_align_fpn_backward`` -> C ABI); nothing here is a different code path.
"""
import numpy as np
import torch

from .. import _lib as L
from .roi_align import roi_align_fpn_backward, roi_align_fpn_forward


def _copy2d(dst, src, rows, width_elems, dst_pitch_elems, src_pitch_elems, h2d, stream):
    L.call("mxd_copy2d_async", dst.data_ptr(), dst_pitch_elems * 4, src.data_ptr(), src_pitch_elems * 4, width_elems * 4,
           rows, 1 if h2d else 2, L.c_void_p(stream.cuda_stream))


class HostRoIStage:
    """Software pipeline  H2D | RoIAlign fwd+bwd | D2H  over ``depth`` device slots; one unit = one image x one
    slice of C / channel_split channels.

    feat_shapes: per level (N, C, H, W) of the HOST feature / gradient tensors (pinned memory).
    RoIs must be grouped by image (non-decreasing batch index), as the RoI samplers emit them.
    """

    def __init__(self, feat_shapes, max_rois_per_image, pooled_size, spatial_scales, sample_ratio=2, device="cuda",
                 depth=2, channel_split=1):
        self.dev = torch.device(device)
        self.shapes = [tuple(int(v) for v in s) for s in feat_shapes]
        self.pooled = (int(pooled_size[0]), int(pooled_size[1]))
        self.scales = [float(s) for s in spatial_scales]
        self.sr = int(sample_ratio)
        self.depth = int(depth)
        C = self.shapes[0][1]
        self.split = int(channel_split)
        if C % self.split:
            raise ValueError("HostRoIStage: channel_split %d does not divide C=%d" % (self.split, C))
        Cs = C // self.split
        self.C, self.Cs = C, Cs
        R = int(max_rois_per_image)
        mk = lambda *s: torch.empty(s, dtype=torch.float32, device=self.dev)   # noqa: E731
        self.slots = [dict(feats=[mk(1, Cs, s[2], s[3]) for s in self.shapes], grads=[mk(1, Cs, s[2], s[3]) for s in self.shapes],
                           rois=mk(R, 5), gout=mk(R, Cs, *self.pooled), out=mk(R, Cs, *self.pooled))
                      for _ in range(self.depth)]
        self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(self.dev) for _ in range(3))
        self.max_rois = R
        # image-local copy of the RoI table (batch index 0): pinned once, cudaHostAlloc per step would serialise the streams
        self._rois0 = torch.empty((self.shapes[0][0] * R, 5), dtype=torch.float32).pin_memory()

    def forward_backward(self, feats_h, rois_h, grad_out_h, out_h, grads_h):
        """Enqueues the whole step and returns an event; ``event.synchronize()`` (or ``torch.cuda.synchronize()``)
        completes it.  out_h (R,C,PH,PW) and grads_h[l] (N,C,H,W) are pinned host destinations."""
        b = rois_h[:, 0].numpy().astype(np.int64)
        if b.size and np.any(np.diff(b) < 0):
            raise ValueError("HostRoIStage: RoIs must be grouped by image (non-decreasing batch index)")
        N = self.shapes[0][0]
        if b.size and (b.min() < 0 or b.max() >= N):
            raise ValueError("HostRoIStage: batch indices must lie in [0, %d) (got %d .. %d)" % (N, b.min(), b.max()))
        counts = np.bincount(b, minlength=N)[:N] if b.size else np.zeros(N, np.int64)
        if counts.max(initial=0) > self.max_rois:
            raise ValueError("HostRoIStage: %d RoIs on one image, capacity %d" % (counts.max(), self.max_rois))
        starts = np.concatenate([[0], np.cumsum(counts)])
        if getattr(self, "_last_in", None) is not None:
            self._last_in.synchronize()                  # the previous step's H2D copies have read the pinned table
        rois0 = self._rois0[: rois_h.shape[0]]
        rois0.copy_(rois_h)
        rois0[:, 0] = 0                                  # every slot holds one image
        bins = self.pooled[0] * self.pooled[1]
        C, Cs = self.C, self.Cs
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.s_in, self.s_cmp, self.s_out):
            s.wait_stream(cur)
        ev_in = [None] * self.depth; ev_cmp = [None] * self.depth; ev_out = [None] * self.depth
        unit = 0
        for i in range(N):
            r0, r1 = int(starts[i]), int(starts[i + 1])
            n = r1 - r0
            for q in range(self.split):
                c0 = q * Cs
                k = unit % self.depth
                unit += 1
                sl = self.slots[k]
                with torch.cuda.stream(self.s_in):
                    if ev_cmp[k] is not None:
                        self.s_in.wait_event(ev_cmp[k])       # the slot's inputs were consumed
                    for l, f in enumerate(feats_h):
                        sl["feats"][l].copy_(f[i:i + 1, c0:c0 + Cs], non_blocking=True)     # contiguous slice
                    if n:
                        sl["rois"][:n].copy_(rois0[r0:r1], non_blocking=True)
                        if self.split == 1:
                            sl["gout"][:n].copy_(grad_out_h[r0:r1], non_blocking=True)
                        else:                                  # (n, Cs, PH, PW) out of (n, C, PH, PW): 2-D copy
                            _copy2d(sl["gout"], grad_out_h[r0, c0], n, Cs * bins, Cs * bins, C * bins, True, self.s_in)
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
                        if self.split == 1:
                            out_h[r0:r1].copy_(sl["out"][:n], non_blocking=True)
                        else:
                            _copy2d(out_h[r0, c0], sl["out"], n, Cs * bins, C * bins, Cs * bins, False, self.s_out)
                    for l, g in enumerate(grads_h):
                        g[i:i + 1, c0:c0 + Cs].copy_(sl["grads"][l], non_blocking=True)
                    ev_out[k] = torch.cuda.Event(); ev_out[k].record(self.s_out)
        self._last_in = torch.cuda.Event(); self._last_in.record(self.s_in)
        for s in (self.s_in, self.s_cmp, self.s_out):
            cur.wait_stream(s)
        done = torch.cuda.Event(); done.record(cur)
        return done
