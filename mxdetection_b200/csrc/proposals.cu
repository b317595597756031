// RPN proposal stage (Spec H, row H1): per (image, level) pre-NMS top-k ->
// anchor regeneration + delta decode + clip + min-size -> NMS -> per-level cap
// -> concat -> per-image top max_num.
//
// Module role: mxdetection/models/rpn_heads (/root/reference/README.md:28);
// RPNHead.get_proposals of mmdet 0.5 / mx.nd.contrib.MultiProposal.
//
// Four launches for the whole batch, no host round trip (the reference style is a
// Python loop over images x levels with a D2H mask reduce):
//   topk_segment (+decode)  ->  nms_mask  ->  nms_resolve  ->  rpn_collect
#include <algorithm>
#include "internal.h"

namespace mxd {

typedef unsigned long long u64;
constexpr int kCollectThreads = 1024;
constexpr int kCollectCap = 56 * 1024;     // concatenated kept scores held in shared memory (224 KB of the 227 KB)

struct CollectArgs {
  const float4* boxes;   // (S,kmax)
  const float* vals;     // (S,kmax)
  const int* keep;       // (S,keep_stride) positions in sorted order
  const int* keep_cnt;   // (S)
  int L, kmax, keep_stride, max_num;
  float* out;            // (B,max_num,5)
  int* num_valid;        // (B)
};

// The per-level keep lists are already in (score desc, index asc) order, so the per-image top max_num
// needs no sort: the global rank of candidate (level l, position p) is p plus, for every other level,
// the number of its candidates that precede it in the total order (score desc, concat index asc) -
// a binary search per level in shared memory.  Rows with rank < max_num are written straight to
// out[rank]; bit-identical to a stable sort of the concatenation.
__global__ void __launch_bounds__(kCollectThreads, 1) rpn_collect_kernel(CollectArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sc = reinterpret_cast<float*>(smem_raw);     // concatenated kept scores (<= kCollectCap)
  __shared__ int s_off[MXD_MAX_LEVELS + 1];
  // grid (B, parts): every CTA loads the image's kept scores (the binary searches need all lists) and ranks
  // a 1/parts share of the candidates - the ranking is a chain of dependent shared-memory reads
  const int b = blockIdx.x, tid = threadIdx.x;
  const int part = blockIdx.y, nparts = gridDim.y;
  if (tid == 0) {
    int acc = 0;
    for (int l = 0; l < a.L; ++l) {
      s_off[l] = acc;
      acc += min(a.keep_cnt[b * a.L + l], a.keep_stride);
    }
    s_off[a.L] = acc;
  }
  __syncthreads();
  const int total = s_off[a.L];
  float* out = a.out + (size_t)b * a.max_num * 5;
  const int nout = min(total, a.max_num);
  for (int l = 0; l < a.L; ++l) {
    const int s = b * a.L + l, cnt = s_off[l + 1] - s_off[l];
    for (int j = tid; j < cnt; j += kCollectThreads)
      sc[s_off[l] + j] = a.vals[(size_t)s * a.kmax + a.keep[(size_t)s * a.keep_stride + j]] + 0.0f;   // -0 -> +0
  }
  if (part == 0) {
    for (int r = nout + tid; r < a.max_num; r += kCollectThreads) {
      float* o = out + (size_t)r * 5;
      o[0] = o[1] = o[2] = o[3] = o[4] = 0.0f;
    }
  }
  __syncthreads();
  for (int ci = part * kCollectThreads + tid; ci < total; ci += kCollectThreads * nparts) {
    int l = 0;
    while (l + 1 < a.L && ci >= s_off[l + 1]) ++l;
    const float v = sc[ci];
    int rank = ci - s_off[l];
    if (total > a.max_num) {
      for (int m = 0; m < a.L; ++m) {
        if (m == l) continue;
        // number of entries of list m with score > v (m > l) or >= v (m < l); lists are descending
        int lo = s_off[m], hi = s_off[m + 1];
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          const float x = sc[mid];
          const bool before = (m < l) ? (x >= v) : (x > v);
          if (before) lo = mid + 1; else hi = mid;
        }
        rank += lo - s_off[m];
      }
    } else {
      rank = ci;
    }
    if (rank < a.max_num) {
      const int s = b * a.L + l;
      const int pos = a.keep[(size_t)s * a.keep_stride + (ci - s_off[l])];
      const float4 bx = a.boxes[(size_t)s * a.kmax + pos];
      float* o = out + (size_t)rank * 5;
      o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w; o[4] = a.vals[(size_t)s * a.kmax + pos];
    }
  }
  if (tid == 0 && part == 0) a.num_valid[b] = nout;
}

// Tail of the data-parallel path: per-image proposals -> ONE fixed-capacity buffer that a single
// all-gather moves: row 0 of image b = {count, image id, 0, 0, 0, 0}, rows 1..max_num = {image id | -1 for padding,
// x1, y1, x2, y2, score}.
__global__ void pack_detections_kernel(const float* __restrict__ props, const int* __restrict__ num_valid, int B, int M,
                                       int first_image_id, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * (M + 1)) return;
  const int b = i / (M + 1), r = i - b * (M + 1);
  const int nv = num_valid[b];
  float* o = out + (size_t)i * 6;
  if (r == 0) {
    o[0] = (float)nv; o[1] = (float)(first_image_id + b); o[2] = o[3] = o[4] = o[5] = 0.0f;
    return;
  }
  const float* p = props + ((size_t)b * M + (r - 1)) * 5;
  o[0] = (r - 1 < nv) ? (float)(first_image_id + b) : -1.0f;
  o[1] = p[0]; o[2] = p[1]; o[3] = p[2]; o[4] = p[3]; o[5] = p[4];
}

struct RpnWs {
  int* idx; float* vals; float4* boxes; uint8_t* valid; int* cnt; u64* mask; int* keep; int* keep_cnt;
  void* sortws; size_t sort_bytes;      // chunk-sort scratch when nms_pre exceeds MXD_SORT_CAP
  int kmax, keep_stride;
  size_t bytes;
};

static int rpn_dims(const mxd_rpn_config* c, int* kmax, int* keep_stride, long long* nmax = nullptr) {
  MXD_REQUIRE(c != nullptr, MXD_EINVAL, "null config");
  MXD_REQUIRE(c->num_levels >= 1 && c->num_levels <= MXD_MAX_LEVELS, MXD_EINVAL, "num_levels %d not in [1,%d]",
              c->num_levels, MXD_MAX_LEVELS);
  MXD_REQUIRE(c->num_base >= 1 && c->num_base <= MXD_MAX_BASE_ANCHORS, MXD_EINVAL, "num_base %d not in [1,%d]",
              c->num_base, MXD_MAX_BASE_ANCHORS);
  MXD_REQUIRE(c->max_num >= 1, MXD_EINVAL, "max_num must be >= 1");
  int km = 0;
  for (int l = 0; l < c->num_levels; ++l) {
    MXD_REQUIRE(c->feat_h[l] >= 0 && c->feat_w[l] >= 0, MXD_EINVAL, "bad feature size");
    const long long n = (long long)c->feat_h[l] * c->feat_w[l] * c->num_base;
    MXD_REQUIRE(n < (1ll << 31), MXD_ENOTSUP, "level too large");
    const long long k = (c->nms_pre > 0 && c->nms_pre < n) ? c->nms_pre : n;
    km = k > km ? (int)k : km;
    if (nmax && n > *nmax) *nmax = n;
  }
  if (km < 1) km = 1;
  const int ks = (c->nms_post > 0 && c->nms_post < km) ? c->nms_post : km;
  // the per-image merge ranks every kept candidate against the other levels' lists in shared memory; the stock
  // training config of the lineage (nms_pre = nms_post = max_num = 2000 on 5 levels = 10 000 candidates) fits
  MXD_REQUIRE((long long)ks * c->num_levels <= kCollectCap, MXD_ENOTSUP,
              "num_levels*nms_post = %lld exceeds %d", (long long)ks * c->num_levels, kCollectCap);
  *kmax = km;
  *keep_stride = ks;
  return MXD_OK;
}

static RpnWs carve_rpn(void* base, int S, int kmax, int ks, long long nmax) {
  RpnWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return (char*)base + o; };
  w.idx = (int*)take(sizeof(int) * (size_t)S * kmax);
  w.vals = (float*)take(sizeof(float) * (size_t)S * kmax);
  w.boxes = (float4*)take(sizeof(float4) * (size_t)S * kmax);
  w.valid = (uint8_t*)take((size_t)S * kmax);
  w.cnt = (int*)take(sizeof(int) * (size_t)S);
  w.mask = (u64*)take(sizeof(u64) * nms_mask_words(S, kmax));
  w.keep = (int*)take(sizeof(int) * (size_t)S * ks);
  w.keep_cnt = (int*)take(sizeof(int) * (size_t)S);
  w.sort_bytes = topk_long_workspace_bytes(S, nmax, kmax);
  w.sortws = take(w.sort_bytes);
  w.kmax = kmax; w.keep_stride = ks;
  w.bytes = off;
  return w;
}

}  // namespace mxd

using namespace mxd;

extern "C" {

size_t mxd_rpn_proposals_workspace_bytes(const mxd_rpn_config* cfg, int batch) {
  int km, ks;
  long long nmax = 0;
  if (rpn_dims(cfg, &km, &ks, &nmax) != MXD_OK || batch < 0) return 0;
  return carve_rpn(nullptr, batch * cfg->num_levels, km, ks, nmax).bytes;
}

int mxd_rpn_proposals_dims(const mxd_rpn_config* cfg, int* kmax, int* keep_stride) {
  MXD_REQUIRE(kmax && keep_stride, MXD_EINVAL, "null output");
  return rpn_dims(cfg, kmax, keep_stride);
}

int mxd_rpn_proposals(const DLTensor* const* scores, const DLTensor* const* deltas, const DLTensor* img_shapes,
                      const mxd_rpn_config* cfg, DLTensor* proposals, DLTensor* num_valid, void* workspace,
                      size_t workspace_bytes, void* stream) {
  int dev = -1, rc, km, ks;
  long long nmax = 0;
  if ((rc = rpn_dims(cfg, &km, &ks, &nmax))) return rc;
  MXD_REQUIRE(scores && deltas, MXD_EINVAL, "null level tables");
  const int L = cfg->num_levels;
  int B = -1;
  TopkParams p = {};
  for (int l = 0; l < L; ++l) {
    if ((rc = check_tensor(scores[l], "scores[l]", F32, 2, 2, &dev))) return rc;
    if ((rc = check_tensor(deltas[l], "deltas[l]", F32, 3, 3, &dev))) return rc;
    const long long n = (long long)cfg->feat_h[l] * cfg->feat_w[l] * cfg->num_base;
    if (B < 0) B = (int)scores[l]->shape[0];
    MXD_REQUIRE(scores[l]->shape[0] == B && scores[l]->shape[1] == n, MXD_EINVAL,
                "scores[%d] must be (B=%d, H*W*A=%lld)", l, B, n);
    MXD_REQUIRE(deltas[l]->shape[0] == B && deltas[l]->shape[1] == n && deltas[l]->shape[2] == 4, MXD_EINVAL,
                "deltas[%d] must be (B=%d, H*W*A=%lld, 4)", l, B, n);
    MXD_REQUIRE(((uintptr_t)dptr<float>(deltas[l]) & 15) == 0, MXD_EINVAL, "deltas[%d] must be 16-byte aligned", l);
    p.scores[l] = dptr<float>(scores[l]);
    p.seg_stride[l] = n;
    p.n[l] = (int)n;
    p.k[l] = (int)((cfg->nms_pre > 0 && cfg->nms_pre < n) ? cfg->nms_pre : n);
    p.deltas[l] = dptr<float>(deltas[l]);
    p.feat_w[l] = cfg->feat_w[l] > 0 ? cfg->feat_w[l] : 1;
    p.stride[l] = cfg->stride[l];
    for (int a = 0; a < cfg->num_base; ++a)
      for (int j = 0; j < 4; ++j) p.base[l][a][j] = cfg->base_anchors[l][a][j];
  }
  if ((rc = check_tensor(img_shapes, "img_shapes", I32, 2, 2, &dev))) return rc;
  MXD_REQUIRE(img_shapes->shape[0] == B && img_shapes->shape[1] == 2, MXD_EINVAL, "img_shapes must be (B,2) [h,w]");
  if ((rc = check_tensor(proposals, "proposals", F32, 3, 3, &dev))) return rc;
  MXD_REQUIRE(proposals->shape[0] == B && proposals->shape[1] == cfg->max_num && proposals->shape[2] == 5, MXD_EINVAL,
              "proposals must be (B=%d,max_num=%d,5)", B, cfg->max_num);
  if ((rc = check_tensor(num_valid, "num_valid", I32, 1, 1, &dev))) return rc;
  MXD_REQUIRE(num_valid->shape[0] == B, MXD_EINVAL, "num_valid must be (B)");
  MXD_REQUIRE(cfg->wh_ratio_clip > 0, MXD_EINVAL, "wh_ratio_clip must be > 0");
  if (B == 0) return MXD_OK;
  const int S = B * L;
  RpnWs w = carve_rpn(workspace, S, km, ks, nmax);
  MXD_REQUIRE(workspace && workspace_bytes >= w.bytes, MXD_EWORKSPACE, "workspace %zu < %zu bytes", workspace_bytes,
              w.bytes);
  MXD_REQUIRE(((uintptr_t)workspace & 255) == 0, MXD_EINVAL, "workspace must be 256-byte aligned");
  cudaStream_t st = as_stream(stream);

  p.num_levels = L; p.batch = B; p.elem_stride = 1; p.kmax = km;
  p.valid_thresh = -INFINITY;
  p.out_idx = w.idx; p.out_val = w.vals; p.out_cnt = w.cnt;
  p.out_boxes = w.boxes; p.out_valid = w.valid;
  p.num_base = cfg->num_base;
  p.img_shapes = dptr<int>(img_shapes);
  for (int j = 0; j < 4; ++j) { p.means[j] = cfg->means[j]; p.stds[j] = cfg->stds[j]; }
  p.max_ratio = (float)fabs(log(cfg->wh_ratio_clip));
  p.min_size = cfg->min_bbox_size;
  if ((rc = launch_topk(p, st, w.sortws, w.sort_bytes))) return rc;

  NmsSortedArgs a = {};
  a.boxes = w.boxes; a.valid = w.valid; a.ids = nullptr; a.counts = w.cnt; a.order = nullptr;
  a.S = S; a.stride = km; a.n_max = km; a.thr = cfg->nms_thr; a.delta = cfg->delta;
  a.max_out = ks; a.mask = w.mask; a.keep = w.keep; a.keep_stride = ks; a.keep_cnt = w.keep_cnt;
  if ((rc = launch_nms_sorted(a, st))) return rc;

  CollectArgs c;
  c.boxes = w.boxes; c.vals = w.vals; c.keep = w.keep; c.keep_cnt = w.keep_cnt;
  c.L = L; c.kmax = km; c.keep_stride = ks; c.max_num = cfg->max_num;
  c.out = dptr<float>(proposals); c.num_valid = dptr<int>(num_valid);
  const int smem = std::max(L * ks, 1) * (int)sizeof(float);
  static unsigned long long seen = 0;
  DeviceOnce once_seen(&seen);
  if (once_seen.first())
    MXD_CUDA_OK(cudaFuncSetAttribute(rpn_collect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kCollectCap * (int)sizeof(float)));
  rpn_collect_kernel<<<dim3(B, 4), kCollectThreads, smem, st>>>(c);
  MXD_POST_LAUNCH("rpn_collect");
  return MXD_OK;
}

int mxd_pack_detections(const DLTensor* proposals, const DLTensor* num_valid, int first_image_id, DLTensor* packed,
                        void* stream) {
  int dev = -1, rc;
  if ((rc = check_tensor(proposals, "proposals", F32, 3, 3, &dev))) return rc;
  if ((rc = check_tensor(num_valid, "num_valid", I32, 1, 1, &dev))) return rc;
  if ((rc = check_tensor(packed, "packed", F32, 3, 3, &dev))) return rc;
  const int B = (int)proposals->shape[0], M = (int)proposals->shape[1];
  MXD_REQUIRE(proposals->shape[2] == 5 && num_valid->shape[0] == B, MXD_EINVAL, "proposals (B,M,5) / num_valid (B)");
  MXD_REQUIRE(packed->shape[0] == B && packed->shape[1] == M + 1 && packed->shape[2] == 6, MXD_EINVAL,
              "packed must be (B=%d, M+1=%d, 6)", B, M + 1);
  if (B == 0) return MXD_OK;
  const long long total = (long long)B * (M + 1);
  MXD_REQUIRE(total < (1ll << 31), MXD_ENOTSUP, "too many rows");
  pack_detections_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(
      dptr<float>(proposals), dptr<int>(num_valid), B, M, first_image_id, dptr<float>(packed));
  MXD_POST_LAUNCH("pack_detections");
  return MXD_OK;
}

int mxd_rpn_proposals_stages(const mxd_rpn_config* cfg, int batch, const void* workspace, size_t workspace_bytes,
                             DLTensor* idx, DLTensor* boxes, DLTensor* keep, DLTensor* counts, void* stream) {
  int dev = -1, rc, km, ks;
  long long nmax = 0;
  if ((rc = rpn_dims(cfg, &km, &ks, &nmax))) return rc;
  const int S = batch * cfg->num_levels;
  RpnWs w = carve_rpn(const_cast<void*>(workspace), S, km, ks, nmax);
  MXD_REQUIRE(workspace && workspace_bytes >= w.bytes, MXD_EWORKSPACE, "workspace too small");
  if ((rc = check_tensor(idx, "idx", I32, 3, 3, &dev))) return rc;
  if ((rc = check_tensor(boxes, "boxes", F32, 4, 4, &dev))) return rc;
  if ((rc = check_tensor(keep, "keep", I32, 3, 3, &dev))) return rc;
  if ((rc = check_tensor(counts, "counts", I32, 3, 3, &dev))) return rc;
  MXD_REQUIRE(numel(idx) == (int64_t)S * km && numel(boxes) == (int64_t)S * km * 4 &&
              numel(keep) == (int64_t)S * ks && numel(counts) == (int64_t)S * 2, MXD_EINVAL,
              "stage buffers must be idx(B,L,%d) boxes(B,L,%d,4) keep(B,L,%d) counts(B,L,2)", km, km, ks);
  if (S == 0) return MXD_OK;
  cudaStream_t st = as_stream(stream);
  MXD_CUDA_OK(cudaMemcpyAsync(dptr<int>(idx), w.idx, sizeof(int) * (size_t)S * km, cudaMemcpyDeviceToDevice, st));
  MXD_CUDA_OK(cudaMemcpyAsync(dptr<float>(boxes), w.boxes, sizeof(float4) * (size_t)S * km, cudaMemcpyDeviceToDevice, st));
  MXD_CUDA_OK(cudaMemcpyAsync(dptr<int>(keep), w.keep, sizeof(int) * (size_t)S * ks, cudaMemcpyDeviceToDevice, st));
  MXD_CUDA_OK(cudaMemcpy2DAsync(dptr<int>(counts), 2 * sizeof(int), w.cnt, sizeof(int), sizeof(int), S,
                                cudaMemcpyDeviceToDevice, st));
  MXD_CUDA_OK(cudaMemcpy2DAsync(dptr<int>(counts) + 1, 2 * sizeof(int), w.keep_cnt, sizeof(int), sizeof(int), S,
                                cudaMemcpyDeviceToDevice, st));
  return MXD_OK;
}

}  // extern "C"
