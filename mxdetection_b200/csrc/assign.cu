// IoU matrix and fused max-IoU assigner (Specs D, E; rows D1/D2).
// Module roles: mxdetection/core/bbox + core/anchor
// (/root/reference/README.md:16-17); bbox_overlaps / bbox_assign_wrt_overlaps of
// mmdet 0.5, mx.nd.contrib.box_iou of mxnet 1.3.0.
//
// The G x N overlap matrix is never materialised.  Pass 1: every thread carries four anchors and walks the GTs held
// in shared memory - but only the GTs whose grown window meets the warp's bounding box (32 windows per ballot);
// it tracks max / first argmax, and the per-GT maximum is reduced warp-wide with one REDUX on the orderable uint
// image of the (non-negative) IoU, then CTA-wide in shared memory, then with one global atomicMax per (CTA, GT).
// Pass 2 re-evaluates the same fp32 expression (bit-identical) to apply the `overlaps[g,n] == gt_max[g]`
// low-quality rule - only for GTs whose maximum does not exceed the anchor's own; the last (largest) g wins, as the
// ascending loop of Spec E.
#include "common.cuh"

namespace mxd {

constexpr int kAssignThreads = 256;
constexpr int kGtChunk = 256;

// Spec D IoU: raw (unclamped) areas, inter = 0 when iw<=0 or ih<=0.
__device__ __forceinline__ float iou_spec_d(const float4& a, float area_a, const float4& g, float area_g, float d) {
  const float iw = __fadd_rn(__fsub_rn(fminf(a.z, g.z), fmaxf(a.x, g.x)), d);
  const float ih = __fadd_rn(__fsub_rn(fminf(a.w, g.w), fmaxf(a.y, g.y)), d);
  const float inter = (iw > 0.0f && ih > 0.0f) ? __fmul_rn(iw, ih) : 0.0f;
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_g, area_a), inter));
}

struct AssignArgs {
  const float4* anchors;   // (N)
  const float4* gts;       // (B,G)
  const int* num_gts;      // (B) or null
  const int* gt_labels;    // (B,G) or null
  const uint8_t* flags;    // (N) or (B,N) or null
  int flags_per_image;
  int B, N, G;
  int* assigned;           // (B,N): pass 1 stores argmax, pass 2 the final value
  float* max_ov;           // (B,N)
  int* labels;             // (B,N) or null
  unsigned int* gt_max;    // (B,G) workspace, zero-initialised (bit image of fp32 >= 0)
  float pos, neg, min_pos, delta;
};

// Pass 1.  Every thread carries kAnchorsPerThread anchors (n, n+256, ...: coalesced), so one shared-memory read
// of a GT box, one warp vote and one REDUX serve 128 anchors of the warp instead of 32 - the per-(warp, GT)
// overhead dominated the one-anchor-per-thread version (260 us for 8 x 268 569 anchors x 100 GTs).
constexpr int kAnchorsPerThread = 4;

__global__ void __launch_bounds__(kAssignThreads) assign_pass1_kernel(AssignArgs a) {
  __shared__ float4 s_gt[kGtChunk];
  __shared__ float4 s_rej[kGtChunk];     // GT box grown by delta + 1: an anchor entirely outside it cannot intersect
  __shared__ float s_area[kGtChunk];
  __shared__ unsigned int s_max[kGtChunk];
  constexpr int A = kAnchorsPerThread;
  const int b = blockIdx.y;
  const int n0 = blockIdx.x * (kAssignThreads * A) + threadIdx.x;
  const int G = a.num_gts ? min(max(a.num_gts[b], 0), a.G) : a.G;
  bool in[A], act[A], posa[A];
  float4 me[A];
  float area[A], best[A];
  int arg[A];
#pragma unroll
  for (int q = 0; q < A; ++q) {
    const int n = n0 + q * kAssignThreads;
    in[q] = n < a.N;
    act[q] = in[q];
    if (in[q] && a.flags) act[q] = a.flags[(a.flags_per_image ? (size_t)b * a.N : 0) + n] != 0;
    me[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    area[q] = 0.f;
    if (act[q]) {
      me[q] = a.anchors[n];
      area[q] = box_area_raw(me[q].x, me[q].y, me[q].z, me[q].w, a.delta);
    }
    posa[q] = area[q] > 0.0f;
    // a rejected pair has IoU +0 (both areas positive): start from "GT 0 with IoU 0", which is what the
    // first iteration of the exact loop would leave behind; degenerate anchors keep the exact -inf start
    best[q] = (posa[q] && G > 0) ? 0.0f : -INFINITY;
    arg[q] = 0;
  }
  // bounding box of the warp's active, positive-area anchors; `wall` = some active anchor has a non-positive
  // (or NaN) area, whose 0/0 cases must reach the exact expression for every GT
  const int lane = threadIdx.x & 31;
  float4 wbox = make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY);
  bool wall = false;
#pragma unroll
  for (int q = 0; q < A; ++q) {
    if (!act[q]) continue;
    if (!posa[q]) { wall = true; continue; }
    wbox.x = fminf(wbox.x, me[q].x); wbox.y = fminf(wbox.y, me[q].y);
    wbox.z = fmaxf(wbox.z, me[q].z); wbox.w = fmaxf(wbox.w, me[q].w);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    wbox.x = fminf(wbox.x, __shfl_xor_sync(0xffffffffu, wbox.x, o));
    wbox.y = fminf(wbox.y, __shfl_xor_sync(0xffffffffu, wbox.y, o));
    wbox.z = fmaxf(wbox.z, __shfl_xor_sync(0xffffffffu, wbox.z, o));
    wbox.w = fmaxf(wbox.w, __shfl_xor_sync(0xffffffffu, wbox.w, o));
  }
  wall = __any_sync(0xffffffffu, wall);
  for (int g0 = 0; g0 < G; g0 += kGtChunk) {
    const int gc = min(kGtChunk, G - g0);
    __syncthreads();
    for (int i = threadIdx.x; i < gc; i += kAssignThreads) {
      const float4 g = a.gts[(size_t)b * a.G + g0 + i];
      s_gt[i] = g;
      const float ga = box_area_raw(g.x, g.y, g.z, g.w, a.delta);
      s_area[i] = ga;
      s_max[i] = 0u;
      // quick-reject window (margin delta + 1 dwarfs every rounding of the exact test); a GT with a
      // non-positive area rejects nothing, so 0/0 cases still reach the exact expression
      const float mg = __fadd_rn(a.delta, 1.0f);
      s_rej[i] = (ga > 0.0f) ? make_float4(__fsub_rn(g.x, mg), __fsub_rn(g.y, mg), __fadd_rn(g.z, mg), __fadd_rn(g.w, mg))
                             : make_float4(-INFINITY, -INFINITY, INFINITY, INFINITY);
    }
    __syncthreads();
    // Most (anchor, GT) pairs do not intersect: their IoU is +0 without the IEEE division (when the union
    // is positive) and cannot raise gt_max.  The warp's 128 anchors have one bounding box; 32 GT windows
    // are tested against it per ballot, and only the surviving GTs (ascending, so ties keep the lowest g)
    // run the exact expression.
    for (int j0 = 0; j0 < gc; j0 += 32) {
      const int gi = j0 + lane;
      bool cand = false;
      if (gi < gc) {
        const float4 rj = s_rej[gi];
        cand = wall || !(wbox.z < rj.x || wbox.x > rj.z || wbox.w < rj.y || wbox.y > rj.w);
      }
      unsigned cmask = __ballot_sync(0xffffffffu, cand);
      while (cmask) {
        const int i = j0 + __ffs(cmask) - 1;
        cmask &= cmask - 1;
        const float4 g = s_gt[i];
        const float ga = s_area[i];
        unsigned bits = 0u;
        bool any_hit = false;
#pragma unroll
        for (int q = 0; q < A; ++q) {
          if (!act[q]) continue;
          const float iw = __fadd_rn(__fsub_rn(fminf(me[q].z, g.z), fmaxf(me[q].x, g.x)), a.delta);
          const float ih = __fadd_rn(__fsub_rn(fminf(me[q].w, g.w), fmaxf(me[q].y, g.y)), a.delta);
          const bool hit = iw > 0.0f && ih > 0.0f;
          float iou = 0.0f;
          if (hit || !(__fadd_rn(ga, area[q]) > 0.0f)) iou = iou_spec_d(me[q], area[q], g, ga, a.delta);
          if (iou > best[q]) { best[q] = iou; arg[q] = g0 + i; }   // strict > keeps the lowest g on ties
          // Spec D/E: 0/0 = NaN propagates through both maxima as in numpy / torch (max_ov = NaN at the first NaN g, so
          // the row stays -1; gt_max = NaN, so the low-quality rule skips that GT).  0x7fc00000 orders above +inf as uint.
          const bool isnan_ = iou != iou;
          if (isnan_ && best[q] == best[q]) { best[q] = iou; arg[q] = g0 + i; }
          any_hit = any_hit || hit || isnan_;
          // IoU is in [+0, 1] or -0 or NaN (a hit needs two boxes of positive width and height): uint order = float order
          if (iou > 0.0f) bits = max(bits, __float_as_uint(iou));
          if (isnan_) bits = 0x7fc00000u;
        }
        if (__any_sync(0xffffffffu, any_hit)) {
          const unsigned wmax = __reduce_max_sync(0xffffffffu, bits);
          if (wmax != 0u && lane == 0) atomicMax(&s_max[i], wmax);
        }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < gc; i += kAssignThreads)
      if (s_max[i] != 0u) atomicMax(&a.gt_max[(size_t)b * a.G + g0 + i], s_max[i]);
  }
#pragma unroll
  for (int q = 0; q < A; ++q) {
    if (!in[q]) continue;
    const size_t o = (size_t)b * a.N + n0 + q * kAssignThreads;
    if (!act[q]) {
      a.assigned[o] = -1;
      a.max_ov[o] = 0.0f;
    } else if (G == 0) {
      a.assigned[o] = 0;
      a.max_ov[o] = 0.0f;
    } else {
      a.assigned[o] = arg[q];     // provisional: argmax, finalised in pass 2
      a.max_ov[o] = best[q];
    }
  }
}

__global__ void __launch_bounds__(kAssignThreads) assign_pass2_kernel(AssignArgs a) {
  __shared__ float4 s_gt[kGtChunk];
  __shared__ float s_area[kGtChunk];
  __shared__ float s_gtmax[kGtChunk];
  __shared__ unsigned int s_gmin;        // bit image of the smallest qualifying gt_max of the chunk
  const int b = blockIdx.y;
  const int n = blockIdx.x * kAssignThreads + threadIdx.x;
  const int G = a.num_gts ? min(max(a.num_gts[b], 0), a.G) : a.G;
  const bool in = n < a.N;
  bool act = in && G > 0;
  if (act && a.flags) act = a.flags[(a.flags_per_image ? (size_t)b * a.N : 0) + n] != 0;
  const size_t o = (size_t)b * a.N + (in ? n : 0);
  float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
  float area = 0.f;
  int result = -1;
  float best = 0.0f;
  if (act) {
    me = a.anchors[n];
    area = box_area_raw(me.x, me.y, me.z, me.w, a.delta);
    best = a.max_ov[o];
    const int arg = a.assigned[o];
    if (best >= 0.0f && best < a.neg) result = 0;
    if (best >= a.pos) result = arg + 1;
  }
  for (int g0 = 0; g0 < G; g0 += kGtChunk) {
    const int gc = min(kGtChunk, G - g0);
    __syncthreads();
    if (threadIdx.x == 0) s_gmin = 0x7f800000u;
    __syncthreads();
    for (int i = threadIdx.x; i < gc; i += kAssignThreads) {
      const float4 g = a.gts[(size_t)b * a.G + g0 + i];
      s_gt[i] = g;
      s_area[i] = box_area_raw(g.x, g.y, g.z, g.w, a.delta);
      const float gm = __uint_as_float(a.gt_max[(size_t)b * a.G + g0 + i]);
      s_gtmax[i] = gm;
      if (gm >= a.min_pos) atomicMin(&s_gmin, __float_as_uint(fmaxf(gm, 0.0f)));   // gt_max >= 0: uint order = float order
    }
    __syncthreads();
    // an anchor whose own maximum is below every qualifying gt_max cannot tie with any of them
    if (act && !(best < __uint_as_float(s_gmin))) {
      for (int i = 0; i < gc; ++i) {
        const float gm = s_gtmax[i];
        if (!(gm >= a.min_pos)) continue;
        if (gm > best) continue;   // overlaps[g,n] <= max_g overlaps[.,n] = best < gm: no tie possible (one compare
                                   // prunes almost every pair: few anchors reach any GT's maximum)
        if (gm > 0.0f) {   // a positive maximum can only be matched by an intersecting pair
          const float4 g = s_gt[i];
          const float iw = __fadd_rn(__fsub_rn(fminf(me.z, g.z), fmaxf(me.x, g.x)), a.delta);
          const float ih = __fadd_rn(__fsub_rn(fminf(me.w, g.w), fmaxf(me.y, g.y)), a.delta);
          if (!(iw > 0.0f && ih > 0.0f)) continue;
        }
        if (iou_spec_d(me, area, s_gt[i], s_area[i], a.delta) == gm) result = g0 + i + 1;
      }
    }
  }
  if (in && (act || G > 0)) {
    if (act) {
      a.assigned[o] = result;
      if (a.labels) a.labels[o] = (result > 0 && a.gt_labels) ? a.gt_labels[(size_t)b * a.G + result - 1] : 0;
    } else if (a.labels) {
      a.labels[o] = 0;
    }
  } else if (in && a.labels) {
    a.labels[o] = 0;
  }
}

__global__ void overlaps_kernel(const float4* __restrict__ b1, const float4* __restrict__ b2, int G, int N,
                                float delta, float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int g = blockIdx.y;
  if (n >= N) return;
  const float4 a = b2[n], q = b1[g];
  out[(size_t)g * N + n] = iou_spec_d(a, box_area_raw(a.x, a.y, a.z, a.w, delta), q,
                                      box_area_raw(q.x, q.y, q.z, q.w, delta), delta);
}

}  // namespace mxd

using namespace mxd;

extern "C" {

int mxd_bbox_overlaps(const DLTensor* b1, const DLTensor* b2, DLTensor* out, float delta, void* stream) {
  int dev = -1, rc;
  if ((rc = check_tensor(b1, "b1", F32, 2, 2, &dev))) return rc;
  if ((rc = check_tensor(b2, "b2", F32, 2, 2, &dev))) return rc;
  if ((rc = check_tensor(out, "out", F32, 2, 2, &dev))) return rc;
  MXD_REQUIRE(b1->shape[1] == 4 && b2->shape[1] == 4, MXD_EINVAL, "boxes must be (*,4)");
  const long long G = b1->shape[0], N = b2->shape[0];
  MXD_REQUIRE(out->shape[0] == G && out->shape[1] == N, MXD_EINVAL, "out must be (G,N)");
  MXD_REQUIRE(G <= 65535 && N < (1ll << 31), MXD_ENOTSUP, "G must be <= 65535");
  MXD_REQUIRE((((uintptr_t)dptr<float>(b1) | (uintptr_t)dptr<float>(b2)) & 15) == 0, MXD_EINVAL,
              "boxes must be 16-byte aligned");
  if (G == 0 || N == 0) return MXD_OK;
  dim3 grid(((int)N + 255) / 256, (int)G);
  overlaps_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(dptr<float>(b1)),
                                                        reinterpret_cast<const float4*>(dptr<float>(b2)), (int)G,
                                                        (int)N, delta, dptr<float>(out));
  MXD_POST_LAUNCH("bbox_overlaps");
  return MXD_OK;
}

size_t mxd_max_iou_assign_workspace_bytes(int batch, int num_gt) {
  return align_up(sizeof(unsigned int) * (size_t)(batch > 0 ? batch : 1) * (num_gt > 0 ? num_gt : 1), 256);
}

int mxd_max_iou_assign(const DLTensor* anchors, const DLTensor* gts, const DLTensor* num_gts,
                       const DLTensor* gt_labels, const DLTensor* flags, DLTensor* assigned,
                       DLTensor* max_overlaps, DLTensor* labels, float pos_iou_thr, float neg_iou_thr,
                       float min_pos_iou, float delta, void* workspace, size_t workspace_bytes, void* stream) {
  int dev = -1, rc;
  if ((rc = check_tensor(anchors, "anchors", F32, 2, 2, &dev))) return rc;
  MXD_REQUIRE(anchors->shape[1] == 4, MXD_EINVAL, "anchors must be (N,4)");
  if ((rc = check_tensor(gts, "gts", F32, 2, 3, &dev))) return rc;
  MXD_REQUIRE(gts->shape[gts->ndim - 1] == 4, MXD_EINVAL, "gts must be (B,G,4) or (G,4)");
  const long long N = anchors->shape[0];
  const int B = gts->ndim == 3 ? (int)gts->shape[0] : 1;
  const int G = (int)gts->shape[gts->ndim - 2];
  MXD_REQUIRE(N < (1ll << 31) && B <= 65535, MXD_ENOTSUP, "problem too large");
  if (num_gts) {
    if ((rc = check_tensor(num_gts, "num_gts", I32, 1, 1, &dev))) return rc;
    MXD_REQUIRE(num_gts->shape[0] == B, MXD_EINVAL, "num_gts must be (B)");
  }
  if (gt_labels) {
    if ((rc = check_tensor(gt_labels, "gt_labels", I32, 1, 2, &dev))) return rc;
    MXD_REQUIRE(numel(gt_labels) == (int64_t)B * G, MXD_EINVAL, "gt_labels must be (B,G)");
  }
  int fpi = 0;
  if (flags) {
    if ((rc = check_tensor(flags, "flags", U8, 1, 2, &dev))) return rc;
    MXD_REQUIRE(numel(flags) == N || numel(flags) == (int64_t)B * N, MXD_EINVAL, "flags must be (N) or (B,N)");
    fpi = (flags->ndim == 2 && flags->shape[0] == B && numel(flags) == (int64_t)B * N && B > 1) ? 1 : 0;
  }
  if ((rc = check_tensor(assigned, "assigned", I32, 1, 2, &dev))) return rc;
  if ((rc = check_tensor(max_overlaps, "max_overlaps", F32, 1, 2, &dev))) return rc;
  MXD_REQUIRE(numel(assigned) == (int64_t)B * N && numel(max_overlaps) == (int64_t)B * N, MXD_EINVAL,
              "assigned / max_overlaps must be (B,N)");
  if (labels) {
    if ((rc = check_tensor(labels, "labels", I32, 1, 2, &dev))) return rc;
    MXD_REQUIRE(numel(labels) == (int64_t)B * N, MXD_EINVAL, "labels must be (B,N)");
  }
  MXD_REQUIRE(((uintptr_t)dptr<float>(anchors) & 15) == 0 && (G == 0 || ((uintptr_t)dptr<float>(gts) & 15) == 0),
              MXD_EINVAL, "anchors / gts must be 16-byte aligned");
  const size_t need = mxd_max_iou_assign_workspace_bytes(B, G);
  MXD_REQUIRE(workspace && workspace_bytes >= need, MXD_EWORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, need);
  if (B == 0 || N == 0) return MXD_OK;
  cudaStream_t st = as_stream(stream);
  AssignArgs a;
  a.anchors = reinterpret_cast<const float4*>(dptr<float>(anchors));
  a.gts = reinterpret_cast<const float4*>(dptr<float>(gts));
  a.num_gts = num_gts ? dptr<int>(num_gts) : nullptr;
  a.gt_labels = gt_labels ? dptr<int>(gt_labels) : nullptr;
  a.flags = flags ? dptr<uint8_t>(flags) : nullptr;
  a.flags_per_image = fpi;
  a.B = B; a.N = (int)N; a.G = G;
  a.assigned = dptr<int>(assigned); a.max_ov = dptr<float>(max_overlaps);
  a.labels = labels ? dptr<int>(labels) : nullptr;
  a.gt_max = static_cast<unsigned int*>(workspace);
  a.pos = pos_iou_thr; a.neg = neg_iou_thr; a.min_pos = min_pos_iou; a.delta = delta;
  MXD_CUDA_OK(cudaMemsetAsync(a.gt_max, 0, need, st));
  dim3 grid(((int)N + kAssignThreads - 1) / kAssignThreads, B);
  const int per_cta = kAssignThreads * kAnchorsPerThread;
  dim3 grid1(((int)N + per_cta - 1) / per_cta, B);
  assign_pass1_kernel<<<grid1, kAssignThreads, 0, st>>>(a);
  MXD_POST_LAUNCH("assign_pass1");
  assign_pass2_kernel<<<grid, kAssignThreads, 0, st>>>(a);
  MXD_POST_LAUNCH("assign_pass2");
  return MXD_OK;
}

}  // extern "C"
