// Mask targets and mask paste: mxdetection/core/mask and models/mask_heads
// (/root/reference/README.md:18,30); mask_target / FCNMaskHead.get_seg_masks of mmdet 0.5.
//
//  * mask target = RoIAlign (Spec A, spatial_scale 1) of the matched uint8 GT mask to mask_size x mask_size, then
//    >= 0.5.  The mask bytes are read directly (no fp32 copy of the mask stack), the arithmetic is Spec A's in strict
//    fp32 (round-to-nearest, no FMA contraction, the spec's operation order), so the thresholded target is bit-exact.
//  * mask paste = the mask head's S x S probabilities resized to the detection's integer box with half-pixel bilinear
//    sampling (the convention of the resize the reference calls; restated as Spec N4 in DESIGN.md), > thr,
//    written into an image-sized uint8 canvas.  One pass writes the whole canvas (zeros outside the box).
#include "roi_align.cuh"

namespace mxd {

__global__ void __launch_bounds__(256) mask_target_kernel(const uint8_t* __restrict__ masks, int G, int H, int W,
                                                           const float* __restrict__ props, int pcols,
                                                           const int* __restrict__ inds, int S, int sr, float thr,
                                                           uint8_t* __restrict__ out_u8, float* __restrict__ out_f) {
  const int p = blockIdx.x;
  const int bin = blockIdx.y * blockDim.x + threadIdx.x;
  if (bin >= S * S) return;
  const int ph = bin / S, pw = bin - ph * S;
  const float* r = props + (size_t)p * pcols;
  const int b = inds[p];
  float res = 0.0f;
  if (b >= 0 && b < G) {
    const float rsw = __fmul_rn(r[0], 1.0f), rsh = __fmul_rn(r[1], 1.0f);
    const float rw = fmaxf(__fsub_rn(__fmul_rn(r[2], 1.0f), rsw), 1.0f), rh = fmaxf(__fsub_rn(__fmul_rn(r[3], 1.0f), rsh), 1.0f);
    const float bh = __fdiv_rn(rh, (float)S), bw = __fdiv_rn(rw, (float)S);
    const int gh = sr > 0 ? sr : (int)ceilf(bh), gw = sr > 0 ? sr : (int)ceilf(bw);
    const uint8_t* m = masks + (size_t)b * H * W;
    float acc = 0.0f;
    for (int iy = 0; iy < gh; ++iy) {
      const AxisTap y = axis_tap(rsh, bh, gh, ph, iy, H, W);
      for (int ix = 0; ix < gw; ++ix) {
        const AxisTap x = axis_tap(rsw, bw, gw, pw, ix, W, 1);
        if (!y.valid || !x.valid) continue;
        const float w1 = __fmul_rn(y.h, x.h), w2 = __fmul_rn(y.h, x.l), w3 = __fmul_rn(y.l, x.h), w4 = __fmul_rn(y.l, x.l);
        const float v1 = (float)m[y.lo + x.lo], v2 = (float)m[y.lo + x.hi], v3 = (float)m[y.hi + x.lo], v4 = (float)m[y.hi + x.hi];
        const float val = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, v1), __fmul_rn(w2, v2)), __fmul_rn(w3, v3)), __fmul_rn(w4, v4));
        acc = __fadd_rn(acc, val);
      }
    }
    res = __fdiv_rn(acc, (float)(gh * gw));
  }
  const size_t o = (size_t)p * S * S + bin;
  if (out_u8) out_u8[o] = res >= thr ? 1 : 0;
  if (out_f) out_f[o] = res;
}

// Spec N4 paste: per detection the integer box (x1i, y1i, w, h) =
// (trunc(x1 / scale), trunc(y1 / scale), max(trunc(x2 / scale) - x1i + 1, 1), ...); canvas pixel (x, y) inside it
// samples the S x S map at sx = (x - x1i + 0.5) * (S / w) - 0.5 (clamped to [0, S-1], weights 0 at the borders),
// bilinear in strict fp32: top = m00 * (1 - lx) + m01 * lx, bot likewise, v = top * (1 - ly) + bot * ly; pixel = v > thr.
__global__ void __launch_bounds__(256) paste_masks_kernel(const float* __restrict__ pred, int C, const int* __restrict__ labels,
                                                           const float* __restrict__ boxes, int bcols, int S, int img_h,
                                                           int img_w, float scale, float thr, uint8_t* __restrict__ out) {
  const int n = blockIdx.z;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= img_w || y >= img_h) return;
  const float* bx = boxes + (size_t)n * bcols;
  const int x1 = (int)__fdiv_rn(bx[0], scale), y1 = (int)__fdiv_rn(bx[1], scale);
  const int x2 = (int)__fdiv_rn(bx[2], scale), y2 = (int)__fdiv_rn(bx[3], scale);
  const int w = max(x2 - x1 + 1, 1), h = max(y2 - y1 + 1, 1);
  uint8_t v = 0;
  const int dx = x - x1, dy = y - y1;
  int cls = 0;
  if (labels) cls = labels[n] + 1;                       // mmdet: mask_pred[i, label + 1] (column 0 = background)
  if (dx >= 0 && dx < w && dy >= 0 && dy < h && cls >= 0 && cls < C) {
    const float* m = pred + ((size_t)n * C + cls) * S * S;
    auto axis = [&](int d, int ext, int* i0, int* i1, float* l) {
      float s = __fsub_rn(__fmul_rn(__fadd_rn((float)d, 0.5f), __fdiv_rn((float)S, (float)ext)), 0.5f);
      int lo = (int)floorf(s);
      float fr = __fsub_rn(s, (float)lo);
      if (lo < 0) { lo = 0; fr = 0.0f; }
      if (lo >= S - 1) { lo = S - 1; fr = 0.0f; }
      *i0 = lo; *i1 = min(lo + 1, S - 1); *l = fr;
    };
    int xa, xb, ya, yb; float lx, ly;
    axis(dx, w, &xa, &xb, &lx);
    axis(dy, h, &ya, &yb, &ly);
    const float hx = __fsub_rn(1.0f, lx), hy = __fsub_rn(1.0f, ly);
    const float top = __fadd_rn(__fmul_rn(m[ya * S + xa], hx), __fmul_rn(m[ya * S + xb], lx));
    const float bot = __fadd_rn(__fmul_rn(m[yb * S + xa], hx), __fmul_rn(m[yb * S + xb], lx));
    v = __fadd_rn(__fmul_rn(top, hy), __fmul_rn(bot, ly)) > thr ? 1 : 0;
  }
  out[((size_t)n * img_h + y) * img_w + x] = v;
}

}  // namespace mxd

using namespace mxd;

extern "C" {

int mxd_mask_target(const DLTensor* gt_masks, const DLTensor* proposals, const DLTensor* gt_inds, DLTensor* target,
                    int mask_size, int sample_ratio, float thr, void* stream) {
  int dev = -1, rc;
  if ((rc = check_tensor(gt_masks, "gt_masks", U8, 3, 3, &dev))) return rc;
  if ((rc = check_tensor(proposals, "proposals", F32, 2, 2, &dev))) return rc;
  if ((rc = check_tensor(gt_inds, "gt_inds", I32, 1, 1, &dev))) return rc;
  MXD_REQUIRE(target != nullptr, MXD_EINVAL, "target: null tensor");
  const bool as_u8 = dtype_is(target, U8);
  if ((rc = check_tensor(target, "target", as_u8 ? U8 : F32, 3, 3, &dev))) return rc;
  const int P = (int)proposals->shape[0], pc = (int)proposals->shape[1];
  MXD_REQUIRE(pc >= 4 && gt_inds->shape[0] == P, MXD_EINVAL, "proposals (P,>=4) / gt_inds (P)");
  MXD_REQUIRE(mask_size >= 1 && target->shape[0] == P && target->shape[1] == mask_size && target->shape[2] == mask_size,
              MXD_EINVAL, "target must be (P=%d,%d,%d)", P, mask_size, mask_size);
  const int G = (int)gt_masks->shape[0], H = (int)gt_masks->shape[1], W = (int)gt_masks->shape[2];
  MXD_REQUIRE(G == 0 || (H >= 1 && W >= 1 && (long long)H * W < (1ll << 31)), MXD_EINVAL, "bad mask size");
  if (P == 0) return MXD_OK;
  MXD_REQUIRE(P <= 2147483647 / 1, MXD_ENOTSUP, "too many proposals");
  dim3 grid(P, (mask_size * mask_size + 255) / 256);
  mask_target_kernel<<<grid, 256, 0, as_stream(stream)>>>(dptr<uint8_t>(gt_masks), G, H, W, dptr<float>(proposals), pc,
                                                         dptr<int>(gt_inds), mask_size, sample_ratio, thr,
                                                         as_u8 ? dptr<uint8_t>(target) : nullptr,
                                                         as_u8 ? nullptr : dptr<float>(target));
  MXD_POST_LAUNCH("mask_target");
  return MXD_OK;
}

int mxd_paste_masks(const DLTensor* mask_pred, const DLTensor* labels, const DLTensor* det_bboxes, DLTensor* im_masks,
                    float scale_factor, float thr, void* stream) {
  int dev = -1, rc;
  if ((rc = check_tensor(mask_pred, "mask_pred", F32, 3, 4, &dev))) return rc;
  if ((rc = check_tensor(det_bboxes, "det_bboxes", F32, 2, 2, &dev))) return rc;
  if ((rc = check_tensor(im_masks, "im_masks", U8, 3, 3, &dev))) return rc;
  const int n = (int)mask_pred->shape[0];
  const int C = mask_pred->ndim == 4 ? (int)mask_pred->shape[1] : 1;
  const int S = (int)mask_pred->shape[mask_pred->ndim - 1];
  MXD_REQUIRE(mask_pred->shape[mask_pred->ndim - 2] == S && S >= 1, MXD_EINVAL, "mask_pred must be (n,[C,]S,S)");
  MXD_REQUIRE(det_bboxes->shape[0] == n && det_bboxes->shape[1] >= 4, MXD_EINVAL, "det_bboxes must be (n,>=4)");
  if (labels) {
    if ((rc = check_tensor(labels, "labels", I32, 1, 1, &dev))) return rc;
    MXD_REQUIRE(labels->shape[0] == n, MXD_EINVAL, "labels must be (n)");
  }
  MXD_REQUIRE(im_masks->shape[0] == n, MXD_EINVAL, "im_masks must be (n,img_h,img_w)");
  MXD_REQUIRE(scale_factor > 0.0f, MXD_EINVAL, "scale_factor must be > 0");
  const int ih = (int)im_masks->shape[1], iw = (int)im_masks->shape[2];
  if (n == 0 || ih == 0 || iw == 0) return MXD_OK;
  MXD_REQUIRE(n <= 65535 && (ih + 7) / 8 <= 65535, MXD_ENOTSUP, "too many detections / rows for one launch");
  dim3 grid((iw + 31) / 32, (ih + 7) / 8, n);
  paste_masks_kernel<<<grid, 256, 0, as_stream(stream)>>>(dptr<float>(mask_pred), C, labels ? dptr<int>(labels) : nullptr,
                                                         dptr<float>(det_bboxes), (int)det_bboxes->shape[1], S, ih, iw,
                                                         scale_factor, thr, dptr<uint8_t>(im_masks));
  MXD_POST_LAUNCH("paste_masks");
  return MXD_OK;
}

}  // extern "C"
