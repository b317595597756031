// Host-side launchers shared between translation units of libmxdet_sm100.so.
#pragma once
#include "common.cuh"

namespace mxd {

// ---- stable top-k (+ optional fused anchor regeneration / delta decode) -------
struct TopkParams {
  int num_levels;                       // segments s = b * num_levels + l
  int batch;
  const float* scores[MXD_MAX_LEVELS];  // element i of segment (b,l): scores[l][b*seg_stride[l] + i*elem_stride]
  long long seg_stride[MXD_MAX_LEVELS];
  int elem_stride;
  int n[MXD_MAX_LEVELS];
  int k[MXD_MAX_LEVELS];                // min(topk, n) per level (above MXD_SORT_CAP: chunk-sort + rank-merge path)
  int kmax;                             // row stride of the outputs
  float valid_thresh;                   // rows with score <= valid_thresh are dropped (-inf: none)
  const int* seg_off;                   // optional (device, S+1): ragged segments of ONE score array (num_levels 1):
                                        // segment s = scores[0][seg_off[s] .. seg_off[s+1]), n[0] / k[0] = upper bounds
  int* out_idx;                         // (S,kmax) i32, -1 padded
  float* out_val;                       // (S,kmax) f32 or null
  int* out_cnt;                         // (S) i32 or null: number of real rows
  // fused decode (Spec F on regenerated anchors, Spec C) -- enabled when out_boxes != null
  float4* out_boxes;                    // (S,kmax)
  uint8_t* out_valid;                   // (S,kmax) min-size flag
  const float* deltas[MXD_MAX_LEVELS];  // (B, n_l, 4)
  int feat_w[MXD_MAX_LEVELS];
  float stride[MXD_MAX_LEVELS];
  int num_base;
  float base[MXD_MAX_LEVELS][MXD_MAX_BASE_ANCHORS][4];
  const int* img_shapes;                // (B,2) [h,w]
  float means[4], stds[4];
  float max_ratio;
  float min_size;
};
// long_ws: scratch of topk_long_workspace_bytes(S, max n, max k) bytes, needed only when some k exceeds MXD_SORT_CAP
int launch_topk(const TopkParams& p, cudaStream_t st, void* long_ws = nullptr, size_t long_ws_bytes = 0);
size_t topk_long_workspace_bytes(int S, long long n_max, int k_max);

// ---- NMS over score-sorted segments ---------------------------------------------
struct NmsSortedArgs {
  const float4* boxes;     // [S][stride], score-descending inside each segment
  const uint8_t* valid;    // [S][stride] or null
  const int* ids;          // [S][stride] or null => class-agnostic
  const int* counts;       // [S] rows per segment, or null => n_max
  const int* order;        // [S][stride] or null: keep[] = order[pos] instead of pos
  int S, stride, n_max;
  float thr, delta;
  int max_out;             // <=0: unlimited
  unsigned long long* mask;  // workspace, nms_mask_words(S, n_max) u64
  int* keep;               // [S][keep_stride], -1 padded
  int keep_stride;
  int* keep_cnt;           // [S]
};
size_t nms_mask_words(int S, int n_max);
int launch_nms_sorted(const NmsSortedArgs& a, cudaStream_t st);

// Spec F encode of one (proposal, gt) pair (strict fp32, correctly rounded log).
__device__ __forceinline__ float4 encode_box(float4 a, float4 b, const float* means, const float* stds) {
  const float px = __fmul_rn(__fadd_rn(a.x, a.z), 0.5f), py = __fmul_rn(__fadd_rn(a.y, a.w), 0.5f);
  const float pw = __fadd_rn(__fsub_rn(a.z, a.x), 1.0f), ph = __fadd_rn(__fsub_rn(a.w, a.y), 1.0f);
  const float gx = __fmul_rn(__fadd_rn(b.x, b.z), 0.5f), gy = __fmul_rn(__fadd_rn(b.y, b.w), 0.5f);
  const float gw = __fadd_rn(__fsub_rn(b.z, b.x), 1.0f), gh = __fadd_rn(__fsub_rn(b.w, b.y), 1.0f);
  const float dx = __fdiv_rn(__fsub_rn(gx, px), pw), dy = __fdiv_rn(__fsub_rn(gy, py), ph);
  const float dw = log_cr(__fdiv_rn(gw, pw)), dh = log_cr(__fdiv_rn(gh, ph));
  return make_float4(__fdiv_rn(__fsub_rn(dx, means[0]), stds[0]), __fdiv_rn(__fsub_rn(dy, means[1]), stds[1]),
                     __fdiv_rn(__fsub_rn(dw, means[2]), stds[2]), __fdiv_rn(__fsub_rn(dh, means[3]), stds[3]));
}

// Spec F decode of one box (strict fp32, correctly rounded exp).
__device__ __forceinline__ float4 decode_box(float4 r, float4 dl, const float* means, const float* stds,
                                             float max_ratio, float hmax, float wmax, bool clip) {
  float dx = __fadd_rn(__fmul_rn(dl.x, stds[0]), means[0]);
  float dy = __fadd_rn(__fmul_rn(dl.y, stds[1]), means[1]);
  float dw = __fadd_rn(__fmul_rn(dl.z, stds[2]), means[2]);
  float dh = __fadd_rn(__fmul_rn(dl.w, stds[3]), means[3]);
  dw = fminf(fmaxf(dw, -max_ratio), max_ratio);
  dh = fminf(fmaxf(dh, -max_ratio), max_ratio);
  float px = __fmul_rn(__fadd_rn(r.x, r.z), 0.5f), py = __fmul_rn(__fadd_rn(r.y, r.w), 0.5f);
  float pw = __fadd_rn(__fsub_rn(r.z, r.x), 1.0f), ph = __fadd_rn(__fsub_rn(r.w, r.y), 1.0f);
  float gw = __fmul_rn(pw, exp_cr(dw)), gh = __fmul_rn(ph, exp_cr(dh));
  float gx = __fadd_rn(px, __fmul_rn(pw, dx)), gy = __fadd_rn(py, __fmul_rn(ph, dy));
  float4 o;
  o.x = __fadd_rn(__fsub_rn(gx, __fmul_rn(gw, 0.5f)), 0.5f);
  o.y = __fadd_rn(__fsub_rn(gy, __fmul_rn(gh, 0.5f)), 0.5f);
  o.z = __fsub_rn(__fadd_rn(gx, __fmul_rn(gw, 0.5f)), 0.5f);
  o.w = __fsub_rn(__fadd_rn(gy, __fmul_rn(gh, 0.5f)), 0.5f);
  if (clip) {
    o.x = fminf(fmaxf(o.x, 0.0f), wmax); o.z = fminf(fmaxf(o.z, 0.0f), wmax);
    o.y = fminf(fmaxf(o.y, 0.0f), hmax); o.w = fminf(fmaxf(o.w, 0.0f), hmax);
  }
  return o;
}

// Block-wide bitonic sort of P (power of two) u64 keys in shared memory, DESCENDING.
// Passes with stride <= 32 never leave a 64-key chunk: a warp takes the chunk into registers (two keys per lane:
// c and c + 32), does stride 32 inside the thread and strides 16..1 with shuffles, and writes it back - the first
// six merge sizes (21 passes) and the last six passes of every later size cost one shared-memory round trip
// each instead of one per pass.  Only strides >= 64 go through shared memory with a block barrier (15 passes for
// 2048 keys).  2048 keys on 1024 threads: 32.5 k cycles with every pass in shared memory, see profiles/README.md.
// blockDim.x must be a multiple of 32.
__device__ __forceinline__ void bitonic_sort_desc(unsigned long long* keys, int P) {
  typedef unsigned long long u64;
  if (P < 64) {          // tiny sorts: plain shared-memory passes
    for (int size = 2; size <= P; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
          const int lo = 2 * t - (t & (stride - 1));
          const int hi = lo + stride;
          const bool up = (lo & size) == 0;
          const u64 a = keys[lo], b = keys[hi];
          if ((a < b) == up) { keys[lo] = b; keys[hi] = a; }
        }
        __syncthreads();
      }
    }
    return;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  auto reg_phase = [&](int size_lo, int size_hi) {
    for (int q = warp; q < (P >> 6); q += nw) {
      const int base = q << 6;
      u64 x0 = keys[base + lane], x1 = keys[base + lane + 32];
      for (int size = size_lo; size <= size_hi; size <<= 1) {
        const bool up0 = ((base + lane) & size) == 0, up1 = ((base + lane + 32) & size) == 0;
        if (size >= 64 && (x0 < x1) == up0) { const u64 t = x0; x0 = x1; x1 = t; }   // stride 32 (up0 == up1)
        for (int j = (size >> 1) < 16 ? (size >> 1) : 16; j > 0; j >>= 1) {
          const u64 y0 = __shfl_xor_sync(0xffffffffu, x0, j), y1 = __shfl_xor_sync(0xffffffffu, x1, j);
          const bool is_lo = (lane & j) == 0;
          x0 = ((is_lo == up0) == (x0 > y0)) ? x0 : y0;     // keep the larger when (low slot of an "up" pair) etc.
          x1 = ((is_lo == up1) == (x1 > y1)) ? x1 : y1;
        }
      }
      keys[base + lane] = x0; keys[base + lane + 32] = x1;
    }
  };
  reg_phase(2, 64);
  __syncthreads();
  for (int size = 128; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride >= 64; stride >>= 1) {
      for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;
        const u64 a = keys[lo], b = keys[hi];
        if ((a < b) == up) { keys[lo] = b; keys[hi] = a; }
      }
      __syncthreads();
    }
    reg_phase(size, size);
    __syncthreads();
  }
}

__device__ __forceinline__ int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace mxd
