// Stable top-k by (score desc, index asc) with optional fused anchor
// regeneration + delta decode (Specs B/H ordering rule, Spec C, Spec F).
//
// Serves mxdetection/ops NMS pre-sort and the pre-NMS top-k of
// mxdetection/models/rpn_heads (/root/reference/README.md:24,28).
//
// One CTA per (image, level) segment; everything stays in shared memory.
//   key = orderable(score) << 32 | (n-1-index)     (unique => order is total)
//   n <= CAP          : sort all keys.
//   n >  CAP          : (1) G strided group maxima (G >= k) give a lower bound L
//                           on the k-th largest key - no atomics;
//                       (2) keys >= L are compacted (warp-aggregated) and sorted;
//                       (3) if more than CAP keys pass (adversarial input) an exact
//                           11-bit radix select narrows them to exactly k first.
//
// Segments longer than CAP (the 160 k-anchor FPN level) run on a thread-block CLUSTER of 8 CTAs (4 when the launch
// carries more segments than 8-CTA clusters can be co-resident): CTA r owns the groups [r*G/8, (r+1)*G/8) and
// therefore a strided 1/8 of the scores, the group maxima are all-gathered
// through distributed shared memory, every CTA derives the same bound L, compacts its own share, and the
// survivors are funnelled into CTA 0 for the final sort (one CTA scanning 650 KB twice measured 138 us).
#include <cooperative_groups.h>
#include <stdlib.h>
#include <algorithm>
#include "internal.h"

namespace cg = cooperative_groups;

namespace mxd {

constexpr int kTopkThreads = 1024;
constexpr int kTopkCluster = 8;               // largest cluster; launches with many segments use 4 (see launch_topk)
constexpr int kCap = MXD_SORT_CAP;
constexpr int kRadixBits = 11;
constexpr int kRadixBins = 1 << kRadixBits;

typedef unsigned long long u64;

__device__ __forceinline__ u64 make_key(float s, int i, int n, float valid_thresh) {
  // dropped row: score <= valid_thresh (valid_thresh = -inf keeps everything but NaN)
  if (!(s > valid_thresh) && !(valid_thresh == -INFINITY && s == s)) return 0ull;
  return ((u64)f32_orderable(s) << 32) | (u64)(uint32_t)(n - 1 - i);
}

// Finds the bin that holds the need-th largest key, walking the 2048-bin histogram from the top with the whole
// CTA: thread t owns bins 2047-2t and 2046-2t, a warp scan plus a scan of the 32 warp totals gives every thread
// the count above its bins.  s_state = {bin, keys still needed inside it, keys in it}.  (One warp walking 64 bins
// per lane serially cost ~2 us per radix pass.)
static_assert(kRadixBins == 2 * kTopkThreads, "two bins per thread");
__device__ __forceinline__ void radix_find_bin(const unsigned int* hist, int need, u64* s_state) {
  __shared__ unsigned s_wsum[32];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int b0 = kRadixBins - 1 - 2 * t;
  const unsigned c0 = hist[b0], c1 = hist[b0 - 1];
  const unsigned sum = c0 + c1;
  unsigned incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_wsum[warp] = incl;
  __syncthreads();
  unsigned wincl = s_wsum[lane];
  const unsigned wown = wincl;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned v = __shfl_up_sync(0xffffffffu, wincl, o);
    if (lane >= o) wincl += v;
  }
  const unsigned base = __shfl_sync(0xffffffffu, wincl - wown, warp);
  unsigned run = base + incl - sum;           // keys in the bins above mine
  if (run < (unsigned)need && run + c0 >= (unsigned)need) {
    s_state[0] = (u64)b0; s_state[1] = (u64)(need - run); s_state[2] = (u64)c0;
  } else {
    run += c0;
    if (run < (unsigned)need && run + c1 >= (unsigned)need) {
      s_state[0] = (u64)(b0 - 1); s_state[1] = (u64)(need - run); s_state[2] = (u64)c1;
    }
  }
}

// Histogram vote of one warp: when every active lane holds the same digit (the top bits of the surviving scores
// nearly always agree) ONE atomic carries the count - 32 same-address shared atomics serialise otherwise.
__device__ __forceinline__ void radix_vote(unsigned int* hist, bool act, unsigned d) {
  const unsigned am = __ballot_sync(0xffffffffu, act);
  if (am == 0u) return;
  const int leader = __ffs(am) - 1;
  const unsigned d0 = __shfl_sync(0xffffffffu, d, leader);
  const unsigned same = __ballot_sync(0xffffffffu, act && d == d0);
  if (same == am) {
    if ((int)(threadIdx.x & 31) == leader) atomicAdd(&hist[d0], (unsigned)__popc(am));
  } else if (act) {
    atomicAdd(&hist[d], 1u);
  }
}

// k-th largest of cnt DISTINCT non-zero 64-bit keys held in shared memory (1 <= k <= cnt): MSB-first 11-bit
// radix passes over a shared-memory histogram.  Far fewer instructions than sorting when only the threshold is
// needed (a 4096-key bitonic sort costs ~4 k instructions per warp).  Returns T: exactly k keys are >= T.
__device__ u64 radix_select_smem(const u64* src, int cnt, int k, unsigned int* hist, u64* s_state, int max_passes = 64) {
  u64 prefix = 0, pmask = 0;
  int need = k;
  int pos = 64;
  while (pos > 0 && max_passes-- > 0) {
    const int width = pos < kRadixBits ? pos : kRadixBits;
    const int shift = pos - width;
    const u64 dmask = (1ull << width) - 1;
    for (int i = threadIdx.x; i < kRadixBins; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int i0 = 0; i0 < cnt; i0 += blockDim.x) {          // warp-uniform trip count: radix_vote is collective
      const int i = i0 + threadIdx.x;
      const u64 key = i < cnt ? src[i] : 0ull;
      radix_vote(hist, key != 0 && (key & pmask) == prefix, (unsigned)((key >> shift) & dmask));
    }
    __syncthreads();
    radix_find_bin(hist, need, s_state);
    __syncthreads();
    const u64 digit = s_state[0];
    need = (int)s_state[1];
    const int binsz = (int)s_state[2];
    prefix |= digit << shift;
    pmask |= dmask << shift;
    pos = shift;
    __syncthreads();
    if (binsz == need) break;  // the whole bin is selected
  }
  return prefix;  // max_passes reached: a lower bound (at least k keys are >= prefix)
}

// Lower bound on the k-th largest key of the segment: (a lower bound of) the k-th largest of the G group maxima
// (every group maximum is a key of the segment, so at least k keys are >= it).  Fewer than k non-zero maxima: take every
// valid key (bound 1).
__device__ u64 select_bound(const u64* maxima, int G, int k, unsigned int* hist, u64* s_state) {
  __shared__ int s_nz;
  if (threadIdx.x == 0) s_nz = 0;
  __syncthreads();
  int nz = 0;
  for (int i = threadIdx.x; i < G; i += blockDim.x) nz += maxima[i] != 0;
  nz = __reduce_add_sync(0xffffffffu, nz);
  if ((threadIdx.x & 31) == 0 && nz) atomicAdd(&s_nz, nz);
  __syncthreads();
  if (s_nz < k) return 1ull;
  // a bound, not the exact k-th maximum: two passes fix sign, exponent and 13 mantissa bits of the score - the
  // keys that slip in below the exact value are a 2^-13 relative sliver of the score range
  return radix_select_smem(maxima, G, k, hist, s_state, 2);
}

// Exact selection of the k-th largest key by MSB-first radix passes (slow path).
__device__ u64 radix_select_kth(const float* __restrict__ sc, int estride, int n, int k, float vt,
                                unsigned int* hist, u64* s_state) {
  int nb = 0;
  while ((1ll << nb) < (long long)n) ++nb;
  u64 prefix = 0, pmask = 0;
  int need = k;
  int pos = 64;
  while (pos > 0) {
    int width, shift;
    if (pos > 32) {  // score part: 11, 11, 10 bits
      width = (pos == 64 || pos == 53) ? 11 : 10;
      shift = pos - width;
    } else {
      if (pos == 32) pos = nb;  // skip the always-zero bits above the index width
      if (pos == 0) break;
      width = pos < kRadixBits ? pos : kRadixBits;
      shift = pos - width;
    }
    const u64 dmask = (1ull << width) - 1;
    for (int i = threadIdx.x; i < kRadixBins; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      u64 key = make_key(sc[(size_t)i * estride], i, n, vt);
      if (key != 0 && (key & pmask) == prefix) atomicAdd(&hist[(unsigned)((key >> shift) & dmask)], 1u);
    }
    __syncthreads();
    radix_find_bin(hist, need, s_state);
    __syncthreads();
    const u64 digit = s_state[0];
    need = (int)s_state[1];
    const int binsz = (int)s_state[2];
    prefix |= digit << shift;
    pmask |= dmask << shift;
    pos = shift;
    __syncthreads();
    if (binsz == need) break;  // the whole bin is selected
  }
  return prefix;  // select keys >= prefix (low unprocessed bits are zero)
}

// Geometry of segment s: uniform (b, l) segments of per-level arrays, or ragged segments of one array (seg_off).
struct SegGeom { int b, l, n, k; const float* sc; };
__device__ __forceinline__ SegGeom seg_geom(const TopkParams& p, int s) {
  SegGeom g;
  if (p.seg_off) {
    const int lo = p.seg_off[s], hi = p.seg_off[s + 1];
    g.b = s; g.l = 0; g.n = max(hi - lo, 0); g.k = min(p.k[0], g.n); g.sc = p.scores[0] + lo;
  } else {
    g.b = s / p.num_levels; g.l = s - g.b * p.num_levels;
    g.n = p.n[g.l]; g.k = p.k[g.l]; g.sc = p.scores[g.l] + (size_t)g.b * p.seg_stride[g.l];
  }
  return g;
}

// One output row j of segment s (key 0 = padding): index, value and - when enabled - the regenerated anchor decoded
// with the row's deltas (Spec C + Spec F) and the min-size flag.
__device__ __forceinline__ void emit_row(const TopkParams& p, int s, const SegGeom& g, int j, u64 key) {
  const bool ok = key != 0;
  const int b = g.b, l = g.l, n = g.n;
  const int i = ok ? n - 1 - (int)(uint32_t)key : -1;
  const float* __restrict__ sc = g.sc;
  p.out_idx[(size_t)s * p.kmax + j] = i;
  if (p.out_val) p.out_val[(size_t)s * p.kmax + j] = ok ? sc[(size_t)i * p.elem_stride] : 0.0f;
  if (p.out_boxes) {
    float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
    uint8_t v = 0;
    if (ok) {
      const float hmax = (float)(p.img_shapes[2 * b] - 1), wmax = (float)(p.img_shapes[2 * b + 1] - 1);
      const int A = p.num_base;
      const int a = i % A, cell = i / A;
      const int x = cell % p.feat_w[l], y = cell / p.feat_w[l];
      const float sx = __fmul_rn((float)x, p.stride[l]), sy = __fmul_rn((float)y, p.stride[l]);
      float4 anc;
      anc.x = __fadd_rn(p.base[l][a][0], sx); anc.y = __fadd_rn(p.base[l][a][1], sy);
      anc.z = __fadd_rn(p.base[l][a][2], sx); anc.w = __fadd_rn(p.base[l][a][3], sy);
      const float4 dl = reinterpret_cast<const float4*>(p.deltas[l])[(size_t)b * n + i];
      box = decode_box(anc, dl, p.means, p.stds, p.max_ratio, hmax, wmax, true);
      v = 1;
      if (p.min_size > 0.f) {
        const float w = __fadd_rn(__fsub_rn(box.z, box.x), 1.0f), h = __fadd_rn(__fsub_rn(box.w, box.y), 1.0f);
        v = (w >= p.min_size && h >= p.min_size) ? 1 : 0;
      }
    }
    p.out_boxes[(size_t)s * p.kmax + j] = box;
    p.out_valid[(size_t)s * p.kmax + j] = v;
  }
}

template <bool CLUSTER>
__global__ void __launch_bounds__(kTopkThreads, 1) topk_segment_kernel(const __grid_constant__ TopkParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  u64* keys = reinterpret_cast<u64*>(smem_raw);  // kCap entries (+ kCap more in cluster launches: the funnel of CTA 0)
  __shared__ unsigned int s_hist[kRadixBins];
  __shared__ u64 s_state[4];
  __shared__ int s_count;
  __shared__ int s_cnts[kTopkCluster];

  const int rank = CLUSTER ? (int)cg::this_cluster().block_rank() : 0;
  const int csz = CLUSTER ? (int)cg::this_cluster().num_blocks() : 1;     // 8 or 4 CTAs
  const int s = CLUSTER ? blockIdx.x / csz : blockIdx.x;
  const SegGeom sg = seg_geom(p, s);
  const int n = sg.n, k = sg.k;
  const int es = p.elem_stride;
  const float vt = p.valid_thresh;
  const float* __restrict__ sc = sg.sc;
  const int tid = threadIdx.x;
  int count;

  bool funnelled = false;
  if (CLUSTER && n > kCap) {
    cg::cluster_group cluster = cg::this_cluster();
    const int G = (k <= 2048) ? 4096 : kCap;
    const int gpc = G / csz;                     // groups owned by this CTA (512 or 1024)
    const int rows = kTopkThreads / gpc;         // threads per group (2 or 1)
    const int gl = tid % gpc, tr = tid / gpc;
    const int nt = (n + G - 1) / G;              // elements of a group: i = t*G + rank*gpc + gl, t < nt
    constexpr int U = 8;
    // (1) maxima of my groups
    u64 m = 0;
    for (int t0 = tr; t0 < nt; t0 += rows * U) {
      float v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long i = (long long)(t0 + u * rows) * G + rank * gpc + gl;
        v[u] = (t0 + u * rows < nt && i < n) ? sc[(size_t)i * es] : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long i = (long long)(t0 + u * rows) * G + rank * gpc + gl;
        const u64 key = (t0 + u * rows < nt && i < n) ? make_key(v[u], (int)i, n, vt) : 0ull;
        m = key > m ? key : m;
      }
    }
    u64* funnel = keys + kCap;                   // scratch here, the candidate funnel of CTA 0 later
    funnel[tid] = m;
    __syncthreads();
    if (tr == 0) {
      for (int r = 1; r < rows; ++r) { const u64 o = funnel[r * gpc + gl]; m = o > m ? o : m; }
      for (int dst = 0; dst < csz; ++dst)                // all-gather: every CTA gets all G maxima
        cluster.map_shared_rank(keys, dst)[rank * gpc + gl] = m;
    }
    cluster.sync();
    u64 L = select_bound(keys, G, k, s_hist, s_state);
    if (tid == 0) s_count = 0;
    __syncthreads();  // everyone is done with the maxima before keys[] is overwritten
    // (2) compaction of my share
    const unsigned lane = tid & 31;
    for (int t0 = tr; t0 < nt; t0 += rows * U) {
      float v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long i = (long long)(t0 + u * rows) * G + rank * gpc + gl;
        v[u] = (t0 + u * rows < nt && i < n) ? sc[(size_t)i * es] : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long i = (long long)(t0 + u * rows) * G + rank * gpc + gl;
        const u64 key = (t0 + u * rows < nt && i < n) ? make_key(v[u], (int)i, n, vt) : 0ull;
        const bool take = key >= L;
        const unsigned bal = __ballot_sync(0xffffffffu, take);
        if (bal) {
          const int leader = __ffs(bal) - 1;
          int b0 = 0;
          if ((int)lane == leader) b0 = atomicAdd(&s_count, __popc(bal));
          b0 = __shfl_sync(0xffffffffu, b0, leader);
          const int pos = b0 + __popc(bal & ((1u << lane) - 1u));
          if (take && pos < kCap) keys[pos] = key;
        }
      }
    }
    __syncthreads();
    const int mine = s_count;
    if (tid == 0) cluster.map_shared_rank(s_cnts, 0)[rank] = mine;
    cluster.sync();
    int off = 0, total = 0;
    {
      const int* c0 = cluster.map_shared_rank(s_cnts, 0);
      for (int q = 0; q < csz; ++q) {
        const int cq = c0[q];
        if (q < rank) off += cq;
        total += cq;
      }
    }
    if (total <= kCap) {          // (3) funnel the survivors into CTA 0
      u64* f0 = cluster.map_shared_rank(funnel, 0);
      for (int j = tid; j < mine; j += kTopkThreads) f0[off + j] = keys[j];
      cluster.sync();
      if (rank != 0) return;
      keys = funnel;
      count = total;
      funnelled = true;
    } else {                      // adversarial input: CTA 0 redoes the segment with the exact radix path
      cluster.sync();             // nobody leaves while its shared memory may still be read
      if (rank != 0) return;
    }
  } else if (rank != 0) {
    return;                       // short segment of a cluster launch: CTA 0 sorts it alone
  }
  if (funnelled) {
    // count / keys are set
  } else if (n <= kCap) {
    for (int i = tid; i < n; i += kTopkThreads) keys[i] = make_key(sc[(size_t)i * es], i, n, vt);
    count = n;
  } else {
    // (1) strided group maxima: element i belongs to group i % G.  Eight independent, coalesced loads are
    // in flight per thread (a dependent one-load-per-iteration loop cost ~50 us on the 200 k level).
    const int G = (k <= 2048) ? 4096 : kCap;
    {
      constexpr int U = 8;
      u64 m[U];
#pragma unroll
      for (int u = 0; u < U; ++u) m[u] = 0;
      for (int base = 0; base < n; base += kTopkThreads * U) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = base + u * kTopkThreads + tid;
          v[u] = i < n ? sc[(size_t)i * es] : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = base + u * kTopkThreads + tid;
          const u64 key = i < n ? make_key(v[u], i, n, vt) : 0ull;
          m[u] = key > m[u] ? key : m[u];
        }
      }
      if (G == 4096) {   // slots u and u+4 are the same group
#pragma unroll
        for (int u = 0; u < 4; ++u) keys[u * kTopkThreads + tid] = m[u] > m[u + 4] ? m[u] : m[u + 4];
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u) keys[u * kTopkThreads + tid] = m[u];
      }
    }
    __syncthreads();
    u64 L = select_bound(keys, G, k, s_hist, s_state);
    if (tid == 0) s_count = 0;
    __syncthreads();  // everyone is done with the maxima before keys[] is overwritten
    // (2) compaction of keys >= L
    const int iters = (n + kTopkThreads - 1) / kTopkThreads;
    const unsigned lane = tid & 31;
    {
      constexpr int U = 8;
      for (int base = 0; base < n; base += kTopkThreads * U) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = base + u * kTopkThreads + tid;
          v[u] = i < n ? sc[(size_t)i * es] : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = base + u * kTopkThreads + tid;
          const u64 key = i < n ? make_key(v[u], i, n, vt) : 0ull;
          const bool take = key >= L;
          const unsigned m = __ballot_sync(0xffffffffu, take);
          if (m) {
            const int leader = __ffs(m) - 1;
            int b0 = 0;
            if ((int)lane == leader) b0 = atomicAdd(&s_count, __popc(m));
            b0 = __shfl_sync(0xffffffffu, b0, leader);
            const int pos = b0 + __popc(m & ((1u << lane) - 1u));
            if (take && pos < kCap) keys[pos] = key;
          }
        }
      }
    }
    __syncthreads();
    count = s_count;
    if (count > kCap) {
      // (3) exact threshold, then re-compact exactly min(k, valid) keys
      __syncthreads();
      const u64 T = radix_select_kth(sc, es, n, k, vt, s_hist, s_state);
      if (tid == 0) s_count = 0;
      __syncthreads();
      for (int it = 0; it < iters; ++it) {
        const int i = it * kTopkThreads + tid;
        u64 key = 0;
        if (i < n) key = make_key(sc[(size_t)i * es], i, n, vt);
        if (key != 0 && key >= T) {
          const int pos = atomicAdd(&s_count, 1);
          if (pos < kCap) keys[pos] = key;
        }
      }
      __syncthreads();
      count = min(s_count, kCap);
    }
  }
  if (count > k) {
    // only the top k are emitted: find their threshold with radix passes and sort k keys instead of `count`
    // (non-zero keys only: zeros are dropped rows and sort last anyway)
    __syncthreads();
    int nzl = 0;
    for (int i = tid; i < count; i += kTopkThreads) nzl += keys[i] != 0;
    if (tid == 0) s_count = 0;
    __syncthreads();
    nzl = __reduce_add_sync(0xffffffffu, nzl);
    if ((tid & 31) == 0 && nzl) atomicAdd(&s_count, nzl);
    __syncthreads();
    const int nzc = s_count;
    __syncthreads();
    if (nzc > k) {
      const u64 T = radix_select_smem(keys, count, k, s_hist, s_state);
      u64 mine[kCap / kTopkThreads];
      int nm = 0;
#pragma unroll
      for (int u = 0; u < kCap / kTopkThreads; ++u) {
        const int i = u * kTopkThreads + tid;
        mine[u] = (i < count) ? keys[i] : 0ull;
      }
      if (tid == 0) s_count = 0;
      __syncthreads();          // every candidate is in registers: the array can be rewritten in place
#pragma unroll
      for (int u = 0; u < kCap / kTopkThreads; ++u)
        if (mine[u] >= T && mine[u] != 0) keys[atomicAdd(&s_count, 1)] = mine[u];
      (void)nm;
      __syncthreads();
      count = s_count;          // == k
    }
  }
  const int P = max(next_pow2(count), 2);
  for (int i = count + tid; i < P; i += kTopkThreads) keys[i] = 0;
  __syncthreads();
  bitonic_sort_desc(keys, P);

  // ---- emit -------------------------------------------------------------------
  for (int j = tid; j < p.kmax; j += kTopkThreads) emit_row(p, s, sg, j, (j < k && j < P) ? keys[j] : 0ull);
  if (p.out_cnt && tid == 0) {
    // real rows = position of the first zero key (keys are sorted, zeros last)
    int lo = 0, hi = min(k, P);
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (keys[mid] != 0) lo = mid + 1; else hi = mid;
    }
    p.out_cnt[s] = lo;
  }
}

// ---- long path: k above the in-CTA sort capacity ------------------------------------------------------------
// (1) every chunk of kCap keys is sorted by one CTA (bitonic, shared memory) into the workspace; (2) the global rank
// of a key is its position in its own chunk plus, for every other chunk, the number of larger keys there (a binary
// search: keys are unique, so the ranks are a permutation) - rows with rank < k are emitted straight to out[rank].
// O(n * chunks * log) work instead of a global radix sort; this path serves box_nms / MultiProposal calls over more
// than 8192 rows, which the detector configs of the lineage do not reach on the hot path.
__global__ void __launch_bounds__(kTopkThreads, 1) topk_chunk_sort_kernel(const __grid_constant__ TopkParams p, u64* ws,
                                                                           int nchunks) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  u64* keys = reinterpret_cast<u64*>(smem_raw);
  const int c = blockIdx.x, s = blockIdx.y;
  const SegGeom sg = seg_geom(p, s);
  const int n = sg.n;
  const float* __restrict__ sc = sg.sc;
  for (int j = threadIdx.x; j < kCap; j += kTopkThreads) {
    const long long i = (long long)c * kCap + j;
    keys[j] = i < n ? make_key(sc[(size_t)i * p.elem_stride], (int)i, n, p.valid_thresh) : 0ull;
  }
  __syncthreads();
  bitonic_sort_desc(keys, kCap);
  u64* o = ws + ((size_t)s * nchunks + c) * kCap;
  for (int j = threadIdx.x; j < kCap; j += kTopkThreads) o[j] = keys[j];
}

__global__ void __launch_bounds__(kTopkThreads, 1) topk_merge_rank_kernel(const __grid_constant__ TopkParams p,
                                                                           const u64* __restrict__ ws, int nchunks) {
  __shared__ int s_valid;
  const int c = blockIdx.x, s = blockIdx.y, tid = threadIdx.x;
  const SegGeom sg = seg_geom(p, s);
  const int k = sg.k;
  const u64* seg = ws + (size_t)s * nchunks * kCap;
  if (tid == 0) s_valid = 0;
  __syncthreads();
  // valid rows of the segment: non-zero keys (sorted descending, zeros last) of every chunk
  for (int q = tid; q < nchunks; q += kTopkThreads) {
    const u64* ch = seg + (size_t)q * kCap;
    int lo = 0, hi = kCap;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (ch[mid] != 0) lo = mid + 1; else hi = mid; }
    if (lo) atomicAdd(&s_valid, lo);
  }
  __syncthreads();
  const int valid = min(s_valid, k);
  if (c == 0) {
    for (int j = valid + tid; j < p.kmax; j += kTopkThreads) emit_row(p, s, sg, j, 0ull);
    if (p.out_cnt && tid == 0) p.out_cnt[s] = valid;
  }
  const u64* mine = seg + (size_t)c * kCap;
  for (int j = tid; j < kCap; j += kTopkThreads) {
    const u64 key = mine[j];
    if (key == 0) break;                          // zeros are last
    int rank = j;
    for (int q = 0; q < nchunks && rank < k; ++q) {
      if (q == c) continue;
      const u64* ch = seg + (size_t)q * kCap;
      int lo = 0, hi = kCap;                       // number of keys of chunk q larger than mine
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (ch[mid] > key) lo = mid + 1; else hi = mid; }
      rank += lo;
    }
    if (rank < k) emit_row(p, s, sg, rank, key);
  }
}

size_t topk_long_workspace_bytes(int S, long long n_max, int k_max) {
  if (k_max <= kCap && getenv("MXD_TOPK_FORCE_LONG") == nullptr) return 0;
  const long long nchunks = (n_max + kCap - 1) / kCap;
  return align_up(sizeof(u64) * (size_t)S * (size_t)nchunks * kCap, 256);
}

int launch_topk(const TopkParams& p, cudaStream_t st, void* long_ws, size_t long_ws_bytes) {
  const int S = p.batch * p.num_levels;
  if (S == 0) return MXD_OK;
  int kbig_all = 0;
  long long nbig_all = 0;
  for (int l = 0; l < p.num_levels; ++l) {
    MXD_REQUIRE(p.k[l] <= p.kmax && p.n[l] >= 0, MXD_EINVAL, "bad top-k geometry");
    kbig_all = std::max(kbig_all, p.k[l]);
    nbig_all = std::max<long long>(nbig_all, p.n[l]);
  }
  static unsigned long long seen_long = 0;
  if (kbig_all > kCap || getenv("MXD_TOPK_FORCE_LONG") != nullptr) {        // (tests force the long path on short inputs)
    const int nchunks = (int)((nbig_all + kCap - 1) / kCap);
    const size_t need = align_up(sizeof(u64) * (size_t)S * (size_t)nchunks * kCap, 256);
    MXD_REQUIRE(long_ws != nullptr && long_ws_bytes >= need, MXD_EWORKSPACE,
                "top-k of %d rows needs a %zu-byte sort workspace (got %zu)", kbig_all, need, long_ws_bytes);
    MXD_REQUIRE(S <= 65535 && nchunks >= 1, MXD_ENOTSUP, "too many top-k segments");
    const int smem = kCap * (int)sizeof(u64);
    DeviceOnce once_seen_long(&seen_long);
  if (once_seen_long.first())
      MXD_CUDA_OK(cudaFuncSetAttribute(topk_chunk_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    topk_chunk_sort_kernel<<<dim3(nchunks, S), kTopkThreads, smem, st>>>(p, static_cast<u64*>(long_ws), nchunks);
    MXD_POST_LAUNCH("topk_chunk_sort");
    topk_merge_rank_kernel<<<dim3(nchunks, S), kTopkThreads, 0, st>>>(p, static_cast<const u64*>(long_ws), nchunks);
    MXD_POST_LAUNCH("topk_merge_rank");
    return MXD_OK;
  }
  bool big = false;
  for (int l = 0; l < p.num_levels; ++l) big = big || p.n[l] > kCap;
  static unsigned long long seen = 0;
  const int smem = kCap * (int)sizeof(u64);
  DeviceOnce once_seen(&seen);
  if (once_seen.first()) {
    MXD_CUDA_OK(cudaFuncSetAttribute(topk_segment_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    MXD_CUDA_OK(cudaFuncSetAttribute(topk_segment_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * smem));
  }
  if (big && (long long)S * kTopkCluster < (1ll << 31)) {
    // 8 CTAs per segment while all clusters are co-resident (one CTA per SM); with more segments than that a
    // second wave of clusters costs more than the longer per-CTA scans of 4-CTA clusters.  G = 8192 group maxima
    // (k > 2048) need 1024 groups per CTA: 8 CTAs.
    int kbig = 0;
    for (int l = 0; l < p.num_levels; ++l) kbig = std::max(kbig, p.k[l]);
    int sms = 148;
    { int dev = 0; if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    const int csz = (S * kTopkCluster > sms && kbig <= 2048) ? 4 : kTopkCluster;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(S * csz)); cfg.blockDim = dim3(kTopkThreads);
    cfg.dynamicSmemBytes = 2 * smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    MXD_CUDA_OK(cudaLaunchKernelEx(&cfg, topk_segment_kernel<true>, p));
  } else {
    topk_segment_kernel<false><<<S, kTopkThreads, smem, st>>>(p);
  }
  MXD_POST_LAUNCH("topk_segment");
  return MXD_OK;
}

}  // namespace mxd

using namespace mxd;

extern "C" {

size_t mxd_topk_stable_workspace_bytes(int segments, int n, int topk) {
  const int k = (topk > 0 && topk < n) ? topk : n;
  return topk_long_workspace_bytes(segments, n, k);
}

int mxd_topk_stable(const DLTensor* scores, DLTensor* idx, DLTensor* vals, int topk, void* workspace,
                    size_t workspace_bytes, void* stream) {
  int dev = -1, rc;
  if ((rc = check_tensor(scores, "scores", F32, 1, 2, &dev))) return rc;
  const int S = scores->ndim == 2 ? (int)scores->shape[0] : 1;
  const long long n = scores->shape[scores->ndim - 1];
  MXD_REQUIRE(n < (1ll << 31), MXD_ENOTSUP, "segment too long");
  const int k = (topk > 0 && topk < n) ? topk : (int)n;
  if ((rc = check_tensor(idx, "idx", I32, scores->ndim, scores->ndim, &dev))) return rc;
  MXD_REQUIRE(numel(idx) == (int64_t)S * k, MXD_EINVAL, "idx must be (S,k) with k=%d", k);
  if (vals) {
    if ((rc = check_tensor(vals, "vals", F32, scores->ndim, scores->ndim, &dev))) return rc;
    MXD_REQUIRE(numel(vals) == (int64_t)S * k, MXD_EINVAL, "vals must be (S,k) with k=%d", k);
  }
  if (S == 0 || k == 0) return MXD_OK;
  TopkParams p = {};
  p.num_levels = 1; p.batch = S;
  p.scores[0] = dptr<float>(scores); p.seg_stride[0] = n; p.elem_stride = 1;
  p.n[0] = (int)n; p.k[0] = k; p.kmax = k;
  p.valid_thresh = -INFINITY;
  p.out_idx = dptr<int>(idx);
  p.out_val = vals ? dptr<float>(vals) : nullptr;
  return launch_topk(p, as_stream(stream), workspace, workspace_bytes);
}

}  // extern "C"
