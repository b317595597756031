// mx.nd.contrib.MultiProposal / Proposal (mxnet 1.3.0) on the library's top-k / NMS kernels.
//
// Contract source: SURVEY.md 8(a) Spec H "MultiProposal alt-mode" and Spec F "MX13 variant"
// (mxnet src/operator/contrib/multi_proposal.{cc,cu}, named by BASELINE.json north_star; the
// module that would call it is mxdetection/models/rpn_heads, /root/reference/README.md:28).
//
//   cls_prob (N, 2A, H, W)  - foreground scores are channels A .. 2A-1
//   bbox_pred (N, 4A, H, W) - deltas of anchor a are channels 4a .. 4a+3
//   im_info (N, 3)          - [height, width, scale]
//   -> rois (N*post_n, 5) [batch, x1, y1, x2, y2], scores (N*post_n, 1)
//
// Per image: anchors on the (h, w, a) grid (index = (h*W + w)*A + a); decode with the +1 convention and no
// dw/dh clamp; clip to the image; positions at or beyond (int)(height/stride), (int)(width/stride) get score
// -1; boxes smaller than rpn_min_size*scale are grown by min_size/2 on every side and get score -1 (mxnet's
// FilterBox); stable sort by score descending, first pre_n; greedy NMS with the +1 IoU and a strict '>';
// first post_n kept boxes, CYCLICALLY repeated to fill post_n rows.
//
// The NCHW inputs are read in place (strided) by the prepare kernel - no transposed copy is made.
#include "internal.h"

namespace mxd {

typedef unsigned long long u64;

struct MxpArgs {
  const float* cls_prob;
  const float* bbox_pred;
  const float* im_info;
  int B, A, H, W;
  float stride, min_size;
  float base[MXD_MAX_BASE_ANCHORS][4];
  float4* boxes;     // (B, n)
  float* scores;     // (B, n)
};

__global__ void __launch_bounds__(256) mxp_prepare_kernel(MxpArgs a) {
  const int n = a.H * a.W * a.A;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (i >= n) return;
  const int an = i % a.A;
  const int w = (i / a.A) % a.W;
  const int h = i / (a.A * a.W);
  const size_t plane = (size_t)a.H * a.W;
  const size_t pix = (size_t)h * a.W + w;
  const float* info = a.im_info + (size_t)b * 3;
  const float im_h = info[0], im_w = info[1];
  const int real_h = (int)__fdiv_rn(im_h, a.stride), real_w = (int)__fdiv_rn(im_w, a.stride);
  float score = a.cls_prob[((size_t)b * 2 * a.A + a.A + an) * plane + pix];
  const float* dp = a.bbox_pred + ((size_t)b * 4 * a.A + 4 * an) * plane + pix;
  const float dx = dp[0], dy = dp[plane], dw = dp[2 * plane], dh = dp[3 * plane];
  const float sx = __fmul_rn((float)w, a.stride), sy = __fmul_rn((float)h, a.stride);
  const float x1 = __fadd_rn(a.base[an][0], sx), y1 = __fadd_rn(a.base[an][1], sy);
  const float x2 = __fadd_rn(a.base[an][2], sx), y2 = __fadd_rn(a.base[an][3], sy);
  // BBoxTransformInv
  const float bw = __fadd_rn(__fsub_rn(x2, x1), 1.0f), bh = __fadd_rn(__fsub_rn(y2, y1), 1.0f);
  const float cx = __fadd_rn(x1, __fmul_rn(0.5f, __fsub_rn(bw, 1.0f)));
  const float cy = __fadd_rn(y1, __fmul_rn(0.5f, __fsub_rn(bh, 1.0f)));
  const float pcx = __fadd_rn(__fmul_rn(dx, bw), cx), pcy = __fadd_rn(__fmul_rn(dy, bh), cy);
  const float pw = __fmul_rn(exp_cr(dw), bw), ph = __fmul_rn(exp_cr(dh), bh);
  const float hw = __fmul_rn(0.5f, __fsub_rn(pw, 1.0f)), hh = __fmul_rn(0.5f, __fsub_rn(ph, 1.0f));
  float4 o;
  o.x = fmaxf(fminf(__fsub_rn(pcx, hw), __fsub_rn(im_w, 1.0f)), 0.0f);
  o.y = fmaxf(fminf(__fsub_rn(pcy, hh), __fsub_rn(im_h, 1.0f)), 0.0f);
  o.z = fmaxf(fminf(__fadd_rn(pcx, hw), __fsub_rn(im_w, 1.0f)), 0.0f);
  o.w = fmaxf(fminf(__fadd_rn(pcy, hh), __fsub_rn(im_h, 1.0f)), 0.0f);
  if (h >= real_h || w >= real_w) score = -1.0f;
  // FilterBox
  const float ms = __fmul_rn(a.min_size, info[2]);
  const float iw = __fadd_rn(__fsub_rn(o.z, o.x), 1.0f), ih = __fadd_rn(__fsub_rn(o.w, o.y), 1.0f);
  if (iw < ms || ih < ms) {
    const float g = __fmul_rn(ms, 0.5f);
    o.x = __fsub_rn(o.x, g); o.y = __fsub_rn(o.y, g); o.z = __fadd_rn(o.z, g); o.w = __fadd_rn(o.w, g);
    score = -1.0f;
  }
  a.boxes[(size_t)b * n + i] = o;
  a.scores[(size_t)b * n + i] = score;
}

__global__ void __launch_bounds__(256) mxp_gather_kernel(const float4* __restrict__ boxes, const int* __restrict__ idx,
                                                          int n, int k, float4* __restrict__ sorted) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (j >= k) return;
  const int i = idx[(size_t)b * k + j];
  sorted[(size_t)b * k + j] = i >= 0 ? boxes[(size_t)b * n + i] : make_float4(0.f, 0.f, 0.f, 0.f);
}

// rois row (b*post_n + r) = [b, box of keep[r % n_keep]] - the cyclic padding of mxnet's PrepareOutput
__global__ void __launch_bounds__(256) mxp_output_kernel(const float4* __restrict__ sorted, const float* __restrict__ vals,
                                                          const int* __restrict__ keep, const int* __restrict__ keep_cnt,
                                                          int k, int post_n, float* __restrict__ rois,
                                                          float* __restrict__ scores) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (r >= post_n) return;
  const int nk = keep_cnt[b];
  float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
  float sc = 0.0f;
  if (nk > 0) {
    const int pos = keep[(size_t)b * post_n + (r % nk)];
    bx = sorted[(size_t)b * k + pos];
    sc = vals[(size_t)b * k + pos];
  }
  float* o = rois + ((size_t)b * post_n + r) * 5;
  o[0] = (float)b; o[1] = bx.x; o[2] = bx.y; o[3] = bx.z; o[4] = bx.w;
  if (scores) scores[(size_t)b * post_n + r] = sc;
}

struct MxpWs {
  float4* boxes; float* scores; int* idx; float* vals; int* cnt; float4* sorted; u64* mask; int* keep; int* keep_cnt;
  void* sortws; size_t sort_bytes;      // chunk-sort scratch when rpn_pre_nms_top_n exceeds MXD_SORT_CAP
  size_t bytes;
};

static MxpWs carve_mxp(void* base, int B, int n, int k, int post_n) {
  MxpWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return (char*)base + o; };
  w.boxes = (float4*)take(sizeof(float4) * (size_t)B * n);
  w.scores = (float*)take(sizeof(float) * (size_t)B * n);
  w.idx = (int*)take(sizeof(int) * (size_t)B * k);
  w.vals = (float*)take(sizeof(float) * (size_t)B * k);
  w.cnt = (int*)take(sizeof(int) * (size_t)B);
  w.sorted = (float4*)take(sizeof(float4) * (size_t)B * k);
  w.mask = (u64*)take(sizeof(u64) * nms_mask_words(B, k));
  w.keep = (int*)take(sizeof(int) * (size_t)B * post_n);
  w.keep_cnt = (int*)take(sizeof(int) * (size_t)B);
  w.sort_bytes = topk_long_workspace_bytes(B, n, k);
  w.sortws = take(w.sort_bytes);
  w.bytes = off;
  return w;
}

static int mxp_dims(int A, int H, int W, int pre_n, int post_n, int* n, int* k) {
  MXD_REQUIRE(A >= 1 && A <= MXD_MAX_BASE_ANCHORS, MXD_EINVAL, "num_anchors %d not in [1,%d]", A, MXD_MAX_BASE_ANCHORS);
  MXD_REQUIRE(H >= 1 && W >= 1 && (long long)H * W * A < (1ll << 31), MXD_EINVAL, "bad feature size");
  MXD_REQUIRE(post_n >= 1, MXD_EINVAL, "rpn_post_nms_top_n must be >= 1");
  *n = H * W * A;
  *k = (pre_n > 0 && pre_n < *n) ? pre_n : *n;
  return MXD_OK;
}

}  // namespace mxd

using namespace mxd;

extern "C" {

size_t mxd_multi_proposal_workspace_bytes(int batch, int num_anchors, int feat_h, int feat_w, int rpn_pre_nms_top_n,
                                          int rpn_post_nms_top_n) {
  int n, k;
  if (batch < 0 || mxp_dims(num_anchors, feat_h, feat_w, rpn_pre_nms_top_n, rpn_post_nms_top_n, &n, &k) != MXD_OK) return 0;
  return carve_mxp(nullptr, batch, n, k, rpn_post_nms_top_n).bytes;
}

int mxd_multi_proposal(const DLTensor* cls_prob, const DLTensor* bbox_pred, const DLTensor* im_info, DLTensor* rois,
                       DLTensor* scores, const float* base_anchors, int num_anchors, float feature_stride,
                       int rpn_pre_nms_top_n, int rpn_post_nms_top_n, float threshold, float rpn_min_size,
                       void* workspace, size_t workspace_bytes, void* stream) {
  int dev = -1, rc, n, k;
  if ((rc = check_tensor(cls_prob, "cls_prob", F32, 4, 4, &dev))) return rc;
  if ((rc = check_tensor(bbox_pred, "bbox_pred", F32, 4, 4, &dev))) return rc;
  if ((rc = check_tensor(im_info, "im_info", F32, 2, 2, &dev))) return rc;
  if ((rc = check_tensor(rois, "rois", F32, 2, 2, &dev))) return rc;
  if (scores && (rc = check_tensor(scores, "scores", F32, 2, 2, &dev))) return rc;
  MXD_REQUIRE(base_anchors, MXD_EINVAL, "null base_anchors");
  const int B = (int)cls_prob->shape[0], A = num_anchors, H = (int)cls_prob->shape[2], W = (int)cls_prob->shape[3];
  const int post_n = rpn_post_nms_top_n;
  if ((rc = mxp_dims(A, H, W, rpn_pre_nms_top_n, post_n, &n, &k))) return rc;
  MXD_REQUIRE(cls_prob->shape[1] == 2 * A, MXD_EINVAL, "cls_prob must be (N, 2A=%d, H, W)", 2 * A);
  MXD_REQUIRE(bbox_pred->shape[0] == B && bbox_pred->shape[1] == 4 * A && bbox_pred->shape[2] == H &&
              bbox_pred->shape[3] == W, MXD_EINVAL, "bbox_pred must be (N=%d, 4A=%d, %d, %d)", B, 4 * A, H, W);
  MXD_REQUIRE(im_info->shape[0] == B && im_info->shape[1] == 3, MXD_EINVAL, "im_info must be (N,3) [h,w,scale]");
  MXD_REQUIRE(rois->shape[0] == (int64_t)B * post_n && rois->shape[1] == 5, MXD_EINVAL, "rois must be (N*post_n=%d, 5)",
              B * post_n);
  if (scores)
    MXD_REQUIRE(scores->shape[0] == (int64_t)B * post_n && scores->shape[1] == 1, MXD_EINVAL, "scores must be (N*post_n, 1)");
  MXD_REQUIRE(feature_stride > 0, MXD_EINVAL, "feature_stride must be > 0");
  if (B == 0) return MXD_OK;
  MxpWs w = carve_mxp(workspace, B, n, k, post_n);
  MXD_REQUIRE(workspace && workspace_bytes >= w.bytes, MXD_EWORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, w.bytes);
  MXD_REQUIRE(((uintptr_t)workspace & 255) == 0, MXD_EINVAL, "workspace must be 256-byte aligned");
  cudaStream_t st = as_stream(stream);

  MxpArgs a = {};
  a.cls_prob = dptr<float>(cls_prob); a.bbox_pred = dptr<float>(bbox_pred); a.im_info = dptr<float>(im_info);
  a.B = B; a.A = A; a.H = H; a.W = W; a.stride = feature_stride; a.min_size = rpn_min_size;
  for (int i = 0; i < A; ++i)
    for (int j = 0; j < 4; ++j) a.base[i][j] = base_anchors[i * 4 + j];
  a.boxes = w.boxes; a.scores = w.scores;
  mxp_prepare_kernel<<<dim3((n + 255) / 256, B), 256, 0, st>>>(a);
  MXD_POST_LAUNCH("multi_proposal_prepare");

  TopkParams p = {};
  p.num_levels = 1; p.batch = B; p.elem_stride = 1; p.kmax = k;
  p.scores[0] = w.scores; p.seg_stride[0] = n; p.n[0] = n; p.k[0] = k;
  p.valid_thresh = -INFINITY;
  p.out_idx = w.idx; p.out_val = w.vals; p.out_cnt = w.cnt;
  if ((rc = launch_topk(p, st, w.sortws, w.sort_bytes))) return rc;
  mxp_gather_kernel<<<dim3((k + 255) / 256, B), 256, 0, st>>>(w.boxes, w.idx, n, k, w.sorted);
  MXD_POST_LAUNCH("multi_proposal_gather");

  NmsSortedArgs s = {};
  s.boxes = w.sorted; s.valid = nullptr; s.ids = nullptr; s.counts = w.cnt; s.order = nullptr;
  s.S = B; s.stride = k; s.n_max = k; s.thr = threshold; s.delta = 1.0f;
  s.max_out = post_n; s.mask = w.mask; s.keep = w.keep; s.keep_stride = post_n; s.keep_cnt = w.keep_cnt;
  if ((rc = launch_nms_sorted(s, st))) return rc;

  mxp_output_kernel<<<dim3((post_n + 255) / 256, B), 256, 0, st>>>(w.sorted, w.vals, w.keep, w.keep_cnt, k, post_n,
                                                                    dptr<float>(rois), scores ? dptr<float>(scores) : nullptr);
  MXD_POST_LAUNCH("multi_proposal_output");
  return MXD_OK;
}

}  // extern "C"
