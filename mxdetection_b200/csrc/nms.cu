// Hard NMS (Spec B): 64-bit suppression bitmask + on-device greedy resolve.
//
// Contract: mx.nd.contrib.box_nms of mxnet 1.3.0 (module mxdetection/ops,
// /root/reference/README.md:24): stable score-descending order, strict
// `iou > thr`, optional class ids / force_suppress, keep indices in score order.
//
// Kernel 1 (mask): grid (col block, group of row blocks, segment).  Thread r of a row block tests its box against
//   the 64 boxes of the column block held in shared memory and writes one u64 word (only the upper triangle
//   runs); diagonal tiles also write the transposed word.  Band-major layout: a row block's words are contiguous.
// Kernel 2 (scan): one CTA of 16 warps per segment.  Warp 0 resolves each block of 64 boxes with a lane-parallel
//   fixed point on the transposed diagonal words; the other warps fold the kept rows' words into the suppression
//   words of the later blocks one step behind; bands arrive by bulk copies on an mbarrier ring.  Nothing leaves
//   the device (MXNet's MultiProposal copies the mask to the host for this step).
//   nms_resolve_global_kernel (row-major mask read from L2) takes segments too long for the scan kernel's ring.
#include <stdlib.h>
#include "internal.h"
#include "ptx.cuh"

namespace mxd {

typedef unsigned long long u64;

// grid (column block, group of 4 row blocks, segment), 256 threads: the 64 boxes of the column block are staged
// once for four row blocks (64-thread CTAs spent most of their short life on that load).
constexpr int kMaskRowBlocks = 4;

template <bool IDS>
__global__ void __launch_bounds__(64 * kMaskRowBlocks) nms_mask_kernel(NmsSortedArgs a, int W, int bandmajor) {
  const int cb = blockIdx.x, rb0 = blockIdx.y * kMaskRowBlocks, s = blockIdx.z;
  if (cb < rb0) return;                                    // the whole CTA lies below the diagonal
  const int n = a.counts ? min(a.counts[s], a.n_max) : a.n_max;
  if (rb0 * 64 >= n || cb * 64 >= n) return;
  const size_t seg = (size_t)s * a.stride;
  __shared__ float4 sb[64];
  __shared__ float sa[64];
  __shared__ int sid[64];
  const int t = threadIdx.x & 63;
  const int rb = rb0 + (threadIdx.x >> 6);
  if (threadIdx.x < 64) {
    const int c = cb * 64 + t;
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    float ar = 0.0f;
    int id = 0;
    if (c < n) {
      b = a.boxes[seg + c];
      ar = box_area_clamped(b.x, b.y, b.z, b.w, a.delta);
      if (IDS) id = a.ids[seg + c];
    }
    sb[t] = b; sa[t] = ar;
    if (IDS) sid[t] = id;
  }
  __syncthreads();
  const int r = rb * 64 + t;
  const bool live = rb <= cb && r < n;                    // only the upper triangle is ever read
  u64 bits = 0;
  if (live) {
    const float4 me = a.boxes[seg + r];
    const float area = box_area_clamped(me.x, me.y, me.z, me.w, a.delta);
    const int myid = IDS ? a.ids[seg + r] : 0;
    const int ncol = min(64, n - cb * 64);
    const int start = (cb == rb) ? t + 1 : 0;  // strictly later boxes only
    // (a fully unrolled, predicated version of this loop measured 50 us against 40 us: the early exits of
    // box_iou_gt on disjoint pairs are worth more than the loop overhead)
    for (int i = start; i < ncol; ++i) {
      if (IDS && sid[i] != myid) continue;
      if (box_iou_gt(me, area, sb[i], sa[i], a.delta, a.thr)) bits |= 1ull << i;
    }
  }
  if (!bandmajor) {
    if (live) a.mask[((size_t)s * a.n_max + r) * W + cb] = bits;
    return;
  }
  // band-major: row block rb owns W + 1 slots of 64 words - slot rb holds the block's TRANSPOSED diagonal words
  // (bit j of row t's word set when the earlier row j of the same block suppresses row t: what the scan kernel's
  // lane-parallel fixed point needs), slot cb + 1 the words of column block cb >= rb.  Slots rb .. W of a row block
  // are one contiguous run: the scan kernel fetches a band with a single bulk copy; stores here are coalesced.
  u64* M = a.mask + (size_t)s * W * (W + 1) * 64;
  if (live) M[((size_t)rb * (W + 1) + cb + 1) * 64 + t] = bits;
  __shared__ u64 sdiag[64];
  const bool diag = rb == cb;                              // warp-uniform (64 threads per row block)
  if (diag) sdiag[t] = bits;
  __syncthreads();
  if (diag && r < n) {
    u64 tr = 0;
#pragma unroll 8
    for (int j = 0; j < 64; ++j) tr |= ((sdiag[j] >> t) & 1ull) << j;
    M[((size_t)rb * (W + 1) + rb) * 64 + t] = tr;
  }
}

// Resolve kernel for segments too long for the scan kernel's shared-memory ring (> ~9400 boxes): one CTA per
// segment, the row-major mask stays in global memory (L2), only the suppression words live in shared memory.  Per
// block of 64 boxes warp 0 walks the diagonal words (64-step chain), then every thread ORs the kept rows' words into
// the suppression words it owns (column w belongs to thread w mod 1024: no atomics).  O(n^2 / 64) words through L2 -
// the price of a segment that mx.nd.contrib.box_nms would resolve with n sequential launches.
constexpr int kResolveThreads = 1024;

__global__ void __launch_bounds__(kResolveThreads) nms_resolve_global_kernel(NmsSortedArgs a, int W) {
  extern __shared__ u64 remv[];   // [W]
  __shared__ u64 s_keep;
  __shared__ int s_done;
  const int s = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = a.counts ? min(a.counts[s], a.n_max) : a.n_max;
  const int nW = (n + 63) >> 6;
  const size_t seg = (size_t)s * a.stride;
  const u64* __restrict__ mask = a.mask + (size_t)s * a.n_max * W;
  int* keep = a.keep + (size_t)s * a.keep_stride;
  const int cap = (a.max_out > 0) ? min(a.max_out, a.keep_stride) : a.keep_stride;
  // rows past n and rows failing the min-size filter start out suppressed
  for (int w = warp; w < nW; w += kResolveThreads / 32) {
    const int r0 = w * 64 + lane, r1 = r0 + 32;
    const bool dead0 = r0 >= n || (a.valid && !a.valid[seg + r0]);
    const bool dead1 = r1 >= n || (a.valid && !a.valid[seg + r1]);
    const u64 d = (u64)__ballot_sync(0xffffffffu, dead0) | ((u64)__ballot_sync(0xffffffffu, dead1) << 32);
    if (lane == 0) remv[w] = d;
  }
  if (tid == 0) s_done = 0;
  __syncthreads();
  int nkeep = 0;               // tracked by warp 0
  for (int j = 0; j < nW; ++j) {
    if (warp == 0) {
      const int r0 = j * 64 + lane, r1 = r0 + 32;
      u64 cur = remv[j];
      const u64 wlo = (r0 < n) ? mask[(size_t)r0 * W + j] : 0ull;
      const u64 whi = (r1 < n) ? mask[(size_t)r1 * W + j] : 0ull;
      u64 keepbits = 0;
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const u64 wi = __shfl_sync(0xffffffffu, i < 32 ? wlo : whi, i & 31);
        if (!((cur >> i) & 1ull)) {
          keepbits |= 1ull << i;
          cur |= wi;
        }
      }
      int c = __popcll(keepbits);
      bool done = false;
      if (nkeep + c >= cap) {  // trim to the first (cap - nkeep) kept boxes
        int extra = nkeep + c - cap;
        while (extra-- > 0) keepbits &= ~(1ull << (63 - __clzll(keepbits)));
        c = cap - nkeep;
        done = true;
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int bit = lane + 32 * half;
        if ((keepbits >> bit) & 1ull) {
          const int pos = nkeep + __popcll(keepbits & ((1ull << bit) - 1ull));
          const int row = j * 64 + bit;
          keep[pos] = a.order ? a.order[seg + row] : row;
        }
      }
      nkeep += c;
      if (lane == 0) { s_keep = keepbits; s_done = done ? 1 : 0; }
    }
    __syncthreads();
    if (s_done) break;
    u64 kb = s_keep;
    for (int w = j + 1 + tid; w < nW; w += kResolveThreads) {
      u64 acc = 0, bits = kb;
      while (bits) {
        const int r = __ffsll((long long)bits) - 1;
        bits &= bits - 1;
        acc |= mask[(size_t)(j * 64 + r) * W + w];
      }
      if (acc) remv[w] |= acc;
    }
    __syncthreads();
  }
  __syncthreads();
  if (warp == 0) {
    for (int i = nkeep + lane; i < a.keep_stride; i += 32) keep[i] = -1;
    if (lane == 0) a.keep_cnt[s] = nkeep;
  }
}


__device__ __forceinline__ u64 warp_or64(u64 x) {
  const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)x);
  const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(x >> 32));
  return ((u64)hi << 32) | lo;
}

// Scan kernel for the band-major mask: one CTA per segment, ONE barrier per block of 64 boxes, no atomics.
//   * warp 0 (the scanner) resolves block b from Removed = dead rows | P[b] | Q with a LANE-PARALLEL fixed point
//     instead of a 64-step chain: lane i holds T[i], the earlier rows of the block that suppress row i (the
//     transposed diagonal word, written by the mask kernel).  Row i is removed once a row of T[i] is kept and kept
//     once every row of T[i] is removed; each round decides at least the lowest undecided row, so the greedy result
//     is reached exactly, in (dependency depth) rounds of four ballots - a handful for real boxes, 64 at worst.
//     Q = what the rows it just kept suppress in block b+1, folded with two REDUX while the words are in its registers;
//   * warps 1.. (the helpers) run one block behind: during scan(b) they fold the rows kept in block b-1 into P[j] for
//     the columns j >= b+1 (select by keep bit, REDUX, one plain shared-memory OR per column: a column belongs to one
//     warp per step) - P[b+1] is complete one full scan before the scanner reads it;
//   * the band of block b (T words, then columns b .. nW-1: one contiguous run) arrives by ONE bulk copy issued
//     `slots - 2` blocks ahead by a helper thread into a ring of `slots` buffers, completion on an mbarrier;
//   * the keep list is written after the loop from the per-block keep words (a store inside the loop would hold the
//     scanner until its a.order load returned).
// 2000 boxes per segment: 77 us for the atomic resolve kernel above; 41 us for this structure with a sequential
// 64-step chain (40 cycles per step: ptxas schedules the shared-memory loads just in time); 31 us with the fixed
// point; 25 us with the prefetch off the scanner warp; 22 us with one bulk copy per band.
constexpr int kScanThreads = 512;
constexpr int kScanMaxSlots = 8;

__global__ void __launch_bounds__(kScanThreads) nms_scan_kernel(NmsSortedArgs a, int W, int slots) {
  extern __shared__ __align__(128) u64 sm[];   // remv[W] | P[W] | keepw[W] | pad[W] | ring[slots][W + 1][64]
  __shared__ int s_stop;
  __shared__ __align__(8) u64 s_full[kScanMaxSlots];
  u64* remv = sm;
  u64* P = sm + W;
  u64* keepw = sm + 2 * W;
  u64* ring = sm + 4 * W;
  const int s = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kScanThreads / 32;
  const int n = a.counts ? min(a.counts[s], a.n_max) : a.n_max;
  const int nW = (n + 63) >> 6;
  const size_t seg = (size_t)s * a.stride;
  const size_t SL = 64 * (size_t)(W + 1);     // one ring buffer = one row block of the mask
  const u64* __restrict__ M = a.mask + (size_t)s * W * SL;
  int* keep = a.keep + (size_t)s * a.keep_stride;
  const int cap = (a.max_out > 0) ? min(a.max_out, a.keep_stride) : a.keep_stride;
  const int dist = slots - 2;
  if (tid == 0) {
    for (int i = 0; i < slots; ++i) mbar_init(&s_full[i], 1);
    fence_mbar_init();
    s_stop = 0x7fffffff;       // last block to scan (set by the scanner when the output is full)
  }
  // rows past n and rows failing the min-size filter start out suppressed
  for (int w = warp; w < nW; w += kWarps) {
    const int r0 = w * 64 + lane, r1 = r0 + 32;
    const bool dead0 = r0 >= n || (a.valid && !a.valid[seg + r0]);
    const bool dead1 = r1 >= n || (a.valid && !a.valid[seg + r1]);
    const u64 d = (u64)__ballot_sync(0xffffffffu, dead0) | ((u64)__ballot_sync(0xffffffffu, dead1) << 32);
    if (lane == 0) { remv[w] = d; P[w] = 0ull; }
  }
  __syncthreads();
  // one helper thread issues the copies: the scanner's time is the kernel's critical path
  auto prefetch = [&](int b, int slot) {
    if (b < nW && tid == 32) {
      const uint32_t bytes = (uint32_t)(nW - b + 1) * 512u;          // T words + columns b .. nW-1
      mbar_arrive_expect_tx(&s_full[slot], bytes);
      bulk_g2s(ring + (size_t)slot * SL, M + (size_t)b * SL + (size_t)b * 64, bytes, &s_full[slot]);
    }
  };
  for (int b = 0; b < dist; ++b) prefetch(b, b);
  int sb = 0, sp = 0, sf = dist;   // ring slots of band b, band b-1 and the band being prefetched
  uint32_t par = 0;                // phase parity of slot sb
  int nkeep = 0;               // scanner only
  u64 Q = 0;                   // scanner only
  for (int b = 0; b < nW; ++b) {
    mbar_wait(&s_full[sb], par);   // every thread sees band b land (the helpers read it one iteration later)
    __syncthreads();           // keepw[b-1] and the helper steps <= b-2 are complete; band b-2's buffer is free
    if (b > s_stop) {          // written one barrier ago: every thread sees the same value
      for (int q = b + 1; q < min(nW, b + dist); ++q) {     // no bulk copy may outlive the CTA
        if (++sb == slots) { sb = 0; par ^= 1u; }
        mbar_wait(&s_full[sb], par);
      }
      break;
    }
    prefetch(b + dist, sf);
    if (warp == 0) {
      const u64* bb = ring + (size_t)sb * SL;               // [0] T words, [1] column b, [2] column b+1, ...
      u64 R = remv[b] | P[b] | Q;                           // removed so far; rows past n are in remv
      u64 nlo = 0, nhi = 0;
      if (b + 1 < nW) { nlo = bb[128 + lane]; nhi = bb[128 + lane + 32]; }
      const u64 A0 = bb[lane], A1 = bb[lane + 32];
      u64 K = 0;
      while (~(K | R)) {
        const u64 dec = K | R;
        const bool u0 = !((dec >> lane) & 1ull), u1 = !((dec >> (lane + 32)) & 1ull);
        const unsigned k0 = __ballot_sync(0xffffffffu, u0 && (A0 & ~R) == 0ull);
        const unsigned k1 = __ballot_sync(0xffffffffu, u1 && (A1 & ~R) == 0ull);
        const unsigned r0 = __ballot_sync(0xffffffffu, u0 && (A0 & K) != 0ull);
        const unsigned r1 = __ballot_sync(0xffffffffu, u1 && (A1 & K) != 0ull);
        K |= ((u64)k1 << 32) | k0;
        R |= ((u64)r1 << 32) | r0;
      }
      u64 keepbits = K;
      int c = __popcll(keepbits);
      bool done = false;
      if (nkeep + c >= cap) {  // trim to the first (cap - nkeep) kept boxes
        int extra = nkeep + c - cap;
        while (extra-- > 0) keepbits &= ~(1ull << (63 - __clzll(keepbits)));
        c = cap - nkeep;
        done = true;
      }
      if (lane == 0) { keepw[b] = keepbits; if (done) s_stop = b; }
      Q = warp_or64((((keepbits >> lane) & 1ull) ? nlo : 0ull) | (((keepbits >> (lane + 32)) & 1ull) ? nhi : 0ull));
      nkeep += c;
    } else if (b >= 1) {
      const u64 kb = keepw[b - 1];
      if (kb) {
        const u64* pb = ring + (size_t)sp * SL;                   // band b-1: column j sits at (j - b + 2) * 64
        const bool k0 = (kb >> lane) & 1ull, k1 = (kb >> (lane + 32)) & 1ull;
        for (int j = b + warp; j < nW; j += kWarps - 1) {         // columns b+1 ..: one warp per column per step
          const u64* col = pb + (size_t)(j - b + 2) * 64;
          const u64 r = warp_or64((k0 ? col[lane] : 0ull) | (k1 ? col[lane + 32] : 0ull));
          if (lane == 0 && r) P[j] |= r;
        }
      }
    }
    sp = sb;
    if (++sb == slots) { sb = 0; par ^= 1u; }
    if (++sf == slots) sf = 0;
  }
  __syncthreads();
  // outputs: exclusive prefix of the per-block keep counts (warp 0), then every thread places its rows
  const int nB = (s_stop == 0x7fffffff) ? nW : min(nW, s_stop + 1);
  int* pre = reinterpret_cast<int*>(P);
  if (warp == 0) {
    int base = 0;
    for (int w0 = 0; w0 < nB; w0 += 32) {
      const int c = (w0 + lane < nB) ? __popcll(keepw[w0 + lane]) : 0;
      int inc = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
      }
      if (w0 + lane < nB) pre[w0 + lane] = base + inc - c;
      base += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) a.keep_cnt[s] = base;
    for (int i = base + lane; i < a.keep_stride; i += 32) keep[i] = -1;
  }
  __syncthreads();
  for (int r = tid; r < nB * 64; r += kScanThreads) {
    const u64 kb = keepw[r >> 6];
    const int bit = r & 63;
    if ((kb >> bit) & 1ull) keep[pre[r >> 6] + __popcll(kb & ((1ull << bit) - 1ull))] = a.order ? a.order[seg + r] : r;
  }
}

// both layouts fit: row-major n_max x W, band-major W x (W + 1) x 64
size_t nms_mask_words(int S, int n_max) {
  const size_t W = (size_t)(n_max + 63) / 64;
  return (size_t)S * (W + 1) * 64 * W;
}

int launch_nms_sorted(const NmsSortedArgs& a, cudaStream_t st) {
  if (a.S == 0) return MXD_OK;
  const int W = (a.n_max + 63) / 64;
  // scan kernel (band-major mask) with a 3- to 8-buffer band ring when it fits shared memory, else the row-major
  // mask + global-memory resolve kernel (segments above ~9400 boxes)
  constexpr int kResolveSmemMax = 226 * 1024;      // 227 KB per CTA minus the kernels' static shared variables
  const size_t Wz = (size_t)(W > 0 ? W : 1);
  int slots = 0;
  for (int sl = kScanMaxSlots; sl >= 3 && !slots; --sl)     // ring depth: the L2 -> shared latency (~1 us) spans several blocks
    if ((Wz * 4 + (Wz + 1) * 64 * (size_t)sl) * sizeof(u64) <= (size_t)kResolveSmemMax) slots = sl;
  if (getenv("MXD_NMS_FORCE_GLOBAL") != nullptr) slots = 0;          // (tests force the long-segment path)
  if (a.n_max > 0) {
    MXD_REQUIRE(a.S <= 65535 && W <= 65535, MXD_ENOTSUP, "too many NMS segments");
    dim3 grid(W, (W + kMaskRowBlocks - 1) / kMaskRowBlocks, a.S);
    if (a.ids) nms_mask_kernel<true><<<grid, 64 * kMaskRowBlocks, 0, st>>>(a, W, slots ? 1 : 0);
    else nms_mask_kernel<false><<<grid, 64 * kMaskRowBlocks, 0, st>>>(a, W, slots ? 1 : 0);
    MXD_POST_LAUNCH("nms_mask");
  }
  static unsigned long long seen = 0;
  DeviceOnce once_seen(&seen);
  if (once_seen.first()) {
    MXD_CUDA_OK(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kResolveSmemMax));
    MXD_CUDA_OK(cudaFuncSetAttribute(nms_resolve_global_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kResolveSmemMax));
  }
  if (slots) {
    nms_scan_kernel<<<a.S, kScanThreads, (Wz * 4 + (Wz + 1) * 64 * (size_t)slots) * sizeof(u64), st>>>(a, W, slots);
    MXD_POST_LAUNCH("nms_scan");
    return MXD_OK;
  }
  const size_t smem = Wz * sizeof(u64);
  MXD_REQUIRE(smem <= (size_t)kResolveSmemMax, MXD_ENOTSUP, "NMS segment of %d boxes exceeds the resolve kernel's shared memory", a.n_max);
  nms_resolve_global_kernel<<<a.S, kResolveThreads, smem, st>>>(a, W);
  MXD_POST_LAUNCH("nms_resolve_global");
  return MXD_OK;
}

// Gathers boxes (and ids) into score order for the public entry points.
__global__ void nms_gather_kernel(const float* __restrict__ boxes, const int* __restrict__ ids,
                                  const int* __restrict__ order, int k, float4* __restrict__ ob,
                                  int* __restrict__ oid) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= k) return;
  const int i = order[j];
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i >= 0) b = reinterpret_cast<const float4*>(boxes)[i];
  ob[j] = b;
  if (oid) oid[j] = (i >= 0 && ids) ? ids[i] : 0;
}

// Ragged segments: order (S,kmax) holds positions inside the segment; boxes / ids are gathered into score order and the
// positions are turned into GLOBAL row indices (what the keep list reports).
__global__ void nms_gather_seg_kernel(const float* __restrict__ boxes, const int* __restrict__ ids,
                                      const int* __restrict__ seg_off, int* __restrict__ order, int kmax,
                                      float4* __restrict__ ob, int* __restrict__ oid) {
  const int s = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= kmax) return;
  const size_t o = (size_t)s * kmax + j;
  const int i = order[o];
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  int id = 0;
  if (i >= 0) {
    const int gi = seg_off[s] + i;
    b = reinterpret_cast<const float4*>(boxes)[gi];
    if (ids) id = ids[gi];
    order[o] = gi;
  }
  ob[o] = b;
  if (oid) oid[o] = id;
}

// ---- MXNet (B,N,K) tensor form ------------------------------------------------
__global__ void boxnms_gather_kernel(const float* __restrict__ data, const int* __restrict__ order, int N,
                                     int K, int kmax, int coord_start, int id_index, int in_format,
                                     float4* __restrict__ ob, int* __restrict__ oid) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= kmax) return;
  const int i = order[(size_t)b * kmax + j];
  float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
  int id = 0;
  if (i >= 0) {
    const float* row = data + ((size_t)b * N + i) * K;
    bx = make_float4(row[coord_start], row[coord_start + 1], row[coord_start + 2], row[coord_start + 3]);
    if (in_format == 1) {  // center -> corner
      const float hw = __fdiv_rn(bx.z, 2.0f), hh = __fdiv_rn(bx.w, 2.0f);
      bx = make_float4(__fsub_rn(bx.x, hw), __fsub_rn(bx.y, hh), __fadd_rn(bx.x, hw), __fadd_rn(bx.y, hh));
    }
    if (id_index >= 0) id = (int)row[id_index];
  }
  ob[(size_t)b * kmax + j] = bx;
  if (oid) oid[(size_t)b * kmax + j] = id;
}

__global__ void boxnms_write_kernel(const float* __restrict__ data, const int* __restrict__ keep,
                                    const int* __restrict__ keep_cnt, int N, int K, int kmax,
                                    int coord_start, int in_format, int out_format,
                                    float* __restrict__ out, int* __restrict__ index) {
  const int b = blockIdx.y;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= N) return;
  const int cnt = keep_cnt[b];
  float* o = out + ((size_t)b * N + r) * K;
  int src = -1;
  if (r < cnt && r < kmax) src = keep[(size_t)b * kmax + r];
  if (src < 0) {
    for (int c = 0; c < K; ++c) o[c] = -1.0f;
  } else {
    const float* row = data + ((size_t)b * N + src) * K;
    for (int c = 0; c < K; ++c) o[c] = row[c];
    if (in_format != out_format) {
      const float a0 = row[coord_start], a1 = row[coord_start + 1], a2 = row[coord_start + 2], a3 = row[coord_start + 3];
      if (out_format == 0) {  // center -> corner
        const float hw = __fdiv_rn(a2, 2.0f), hh = __fdiv_rn(a3, 2.0f);
        o[coord_start] = __fsub_rn(a0, hw); o[coord_start + 1] = __fsub_rn(a1, hh);
        o[coord_start + 2] = __fadd_rn(a0, hw); o[coord_start + 3] = __fadd_rn(a1, hh);
      } else {  // corner -> center
        const float w = __fsub_rn(a2, a0), h = __fsub_rn(a3, a1);
        o[coord_start] = __fadd_rn(a0, __fdiv_rn(w, 2.0f)); o[coord_start + 1] = __fadd_rn(a1, __fdiv_rn(h, 2.0f));
        o[coord_start + 2] = w; o[coord_start + 3] = h;
      }
    }
  }
  if (index) index[(size_t)b * N + r] = src;
}

__global__ void boxnms_backward_kernel(const float* __restrict__ og, const int* __restrict__ index, int N, int K,
                                       float* __restrict__ ig) {
  const int b = blockIdx.y;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= N) return;
  const int src = index[(size_t)b * N + r];
  if (src < 0) return;
  const float* g = og + ((size_t)b * N + r) * K;
  float* o = ig + ((size_t)b * N + src) * K;
  for (int c = 0; c < K; ++c) o[c] = g[c];
}

// ---- detection post-processing (BBoxHead.get_det_bboxes / multiclass_nms of mmdet 0.5) ----------
struct DetF4 { float v[4]; };

// candidate m = i * (C-1) + (c-1), c = 1..C-1: box (decoded per Spec F when deltas are given, clipped, divided by
// the scale factor), score = cls_score[i,c], id = c-1
__global__ void det_candidates_kernel(const float* __restrict__ boxes, int box_cols, const float* __restrict__ deltas,
                                      int delta_cols, const float* __restrict__ score, int n, int C, DetF4 means,
                                      DetF4 stds, float max_ratio, float hmax, float wmax, int clip, float scale,
                                      float4* __restrict__ ob, float* __restrict__ os, int* __restrict__ oid) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= (long long)n * (C - 1)) return;
  const int i = (int)(m / (C - 1)), c = (int)(m - (long long)i * (C - 1)) + 1;
  const float* bp = boxes + (size_t)i * box_cols + (box_cols == 4 ? 0 : 4 * c);
  float4 b = make_float4(bp[0], bp[1], bp[2], bp[3]);
  if (deltas) {
    const float* dp = deltas + (size_t)i * delta_cols + (delta_cols == 4 ? 0 : 4 * c);
    b = decode_box(b, make_float4(dp[0], dp[1], dp[2], dp[3]), means.v, stds.v, max_ratio, hmax, wmax, clip != 0);
  }
  // BBoxHead.get_det_bboxes rescales in both branches (decoded or rois-only boxes)
  if (scale != 1.0f) b = make_float4(__fdiv_rn(b.x, scale), __fdiv_rn(b.y, scale), __fdiv_rn(b.z, scale), __fdiv_rn(b.w, scale));
  ob[m] = b;
  os[m] = score[(size_t)i * C + c];
  oid[m] = c - 1;
}

__global__ void det_output_kernel(const float4* __restrict__ cb, const float* __restrict__ cs, const int* __restrict__ cid,
                                  const int* __restrict__ keep, int cap, float* __restrict__ dets, int* __restrict__ labels) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cap) return;
  const int m = keep[j];
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  float sc = 0.f;
  int lab = -1;
  if (m >= 0) { b = cb[m]; sc = cs[m]; lab = cid[m]; }
  float* o = dets + (size_t)j * 5;
  o[0] = b.x; o[1] = b.y; o[2] = b.z; o[3] = b.w; o[4] = sc;
  labels[j] = lab;
}

struct NmsWs {
  int* order; float* vals; int* cnt; float4* boxes; int* ids; u64* mask; int* keep; int* keep_cnt;
  void* sortws; size_t sort_bytes;       // chunk-sort scratch, only when kmax exceeds MXD_SORT_CAP
  size_t bytes;
};

static NmsWs carve(void* base, int S, int kmax, long long n = 0) {
  NmsWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return (char*)base + o; };
  w.order = (int*)take(sizeof(int) * (size_t)S * kmax);
  w.vals = (float*)take(sizeof(float) * (size_t)S * kmax);
  w.cnt = (int*)take(sizeof(int) * (size_t)S);
  w.boxes = (float4*)take(sizeof(float4) * (size_t)S * kmax);
  w.ids = (int*)take(sizeof(int) * (size_t)S * kmax);
  w.mask = (u64*)take(sizeof(u64) * nms_mask_words(S, kmax));
  w.keep = (int*)take(sizeof(int) * (size_t)S * kmax);
  w.keep_cnt = (int*)take(sizeof(int) * (size_t)S);
  w.sort_bytes = topk_long_workspace_bytes(S, n, kmax);
  w.sortws = take(w.sort_bytes);
  w.bytes = off;
  return w;
}

static inline int eff_k(long long n, int topk) { return (int)((topk > 0 && topk < n) ? topk : n); }

}  // namespace mxd

using namespace mxd;

extern "C" {

size_t mxd_nms_workspace_bytes(int n, int topk) { return carve(nullptr, 1, eff_k(n, topk), n).bytes; }

int mxd_nms(const DLTensor* boxes, const DLTensor* scores, const DLTensor* ids, DLTensor* keep,
            DLTensor* num_keep, float iou_thr, float delta, int topk, float valid_thresh,
            int force_suppress, int max_out, void* workspace, size_t workspace_bytes, void* stream) {
  int dev = -1, rc;
  if ((rc = check_tensor(boxes, "boxes", F32, 2, 2, &dev))) return rc;
  MXD_REQUIRE(boxes->shape[1] == 4, MXD_EINVAL, "boxes must be (n,4)");
  const long long n = boxes->shape[0];
  if ((rc = check_tensor(scores, "scores", F32, 1, 1, &dev))) return rc;
  MXD_REQUIRE(scores->shape[0] == n, MXD_EINVAL, "scores must be (n)");
  if (ids) {
    if ((rc = check_tensor(ids, "ids", I32, 1, 1, &dev))) return rc;
    MXD_REQUIRE(ids->shape[0] == n, MXD_EINVAL, "ids must be (n)");
  }
  if ((rc = check_tensor(keep, "keep", I32, 1, 1, &dev))) return rc;
  if ((rc = check_tensor(num_keep, "num_keep", I32, 1, 1, &dev))) return rc;
  MXD_REQUIRE(num_keep->shape[0] >= 1, MXD_EINVAL, "num_keep must hold one int32");
  MXD_REQUIRE(((uintptr_t)dptr<float>(boxes) & 15) == 0, MXD_EINVAL, "boxes must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  const int k = eff_k(n, topk);
  const int cap = (int)keep->shape[0];
  if (n == 0 || cap == 0) {
    MXD_CUDA_OK(cudaMemsetAsync(dptr<int>(num_keep), 0, sizeof(int), st));
      return MXD_OK;
  }
  NmsWs w = carve(workspace, 1, k, n);
  MXD_REQUIRE(workspace && workspace_bytes >= w.bytes, MXD_EWORKSPACE, "workspace %zu < %zu bytes",
              workspace_bytes, w.bytes);
  TopkParams p = {};
  p.num_levels = 1; p.batch = 1;
  p.scores[0] = dptr<float>(scores); p.seg_stride[0] = n; p.elem_stride = 1;
  p.n[0] = (int)n; p.k[0] = k; p.kmax = k;
  p.valid_thresh = valid_thresh;
  p.out_idx = w.order; p.out_val = nullptr; p.out_cnt = w.cnt;
  if ((rc = launch_topk(p, st, w.sortws, w.sort_bytes))) return rc;
  const bool class_aware = ids && !force_suppress;
  nms_gather_kernel<<<(k + 255) / 256, 256, 0, st>>>(dptr<float>(boxes), ids ? dptr<int>(ids) : nullptr, w.order,
                                                      k, w.boxes, class_aware ? w.ids : nullptr);
  MXD_POST_LAUNCH("nms_gather");
  NmsSortedArgs a = {};
  a.boxes = w.boxes; a.valid = nullptr; a.ids = class_aware ? w.ids : nullptr; a.counts = w.cnt;
  a.order = w.order; a.S = 1; a.stride = k; a.n_max = k; a.thr = iou_thr; a.delta = delta;
  a.max_out = max_out; a.mask = w.mask;
  // keep tensor may be shorter than k: resolve writes at most keep_stride rows
  a.keep = dptr<int>(keep); a.keep_stride = cap; a.keep_cnt = dptr<int>(num_keep);
  return launch_nms_sorted(a, st);
}

size_t mxd_nms_batched_workspace_bytes(int num_segments, int max_seg_len, int topk) {
  if (num_segments < 0 || max_seg_len < 0) return 0;
  return carve(nullptr, num_segments, eff_k(max_seg_len, topk), max_seg_len).bytes;
}

int mxd_nms_batched(const DLTensor* boxes, const DLTensor* scores, const DLTensor* ids, const DLTensor* seg_offsets,
                    int max_seg_len, DLTensor* keep, DLTensor* num_keep, float iou_thr, float delta, int topk,
                    float valid_thresh, int force_suppress, int max_out, void* workspace, size_t workspace_bytes,
                    void* stream) {
  int dev = -1, rc;
  if ((rc = check_tensor(boxes, "boxes", F32, 2, 2, &dev))) return rc;
  MXD_REQUIRE(boxes->shape[1] == 4, MXD_EINVAL, "boxes must be (n,4)");
  const long long n = boxes->shape[0];
  if ((rc = check_tensor(scores, "scores", F32, 1, 1, &dev))) return rc;
  MXD_REQUIRE(scores->shape[0] == n, MXD_EINVAL, "scores must be (n)");
  if (ids) {
    if ((rc = check_tensor(ids, "ids", I32, 1, 1, &dev))) return rc;
    MXD_REQUIRE(ids->shape[0] == n, MXD_EINVAL, "ids must be (n)");
  }
  if ((rc = check_tensor(seg_offsets, "seg_offsets", I32, 1, 1, &dev))) return rc;
  MXD_REQUIRE(seg_offsets->shape[0] >= 1, MXD_EINVAL, "seg_offsets must be (S+1)");
  const int S = (int)seg_offsets->shape[0] - 1;
  if ((rc = check_tensor(keep, "keep", I32, 2, 2, &dev))) return rc;
  if ((rc = check_tensor(num_keep, "num_keep", I32, 1, 1, &dev))) return rc;
  MXD_REQUIRE(keep->shape[0] == S && num_keep->shape[0] == S, MXD_EINVAL, "keep / num_keep must be (S=%d, cap) / (S)", S);
  MXD_REQUIRE(max_seg_len >= 0 && max_seg_len <= n, MXD_EINVAL, "max_seg_len %d not in [0, n]", max_seg_len);
  MXD_REQUIRE(n == 0 || ((uintptr_t)dptr<float>(boxes) & 15) == 0, MXD_EINVAL, "boxes must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  const int cap = (int)keep->shape[1];
  if (S == 0) return MXD_OK;
  const int k = eff_k(max_seg_len, topk);
  if (k == 0 || cap == 0) {
    MXD_CUDA_OK(cudaMemsetAsync(dptr<int>(num_keep), 0, sizeof(int) * (size_t)S, st));
    if (cap) MXD_CUDA_OK(cudaMemsetAsync(dptr<int>(keep), 0xff, sizeof(int) * (size_t)S * cap, st));
    return MXD_OK;
  }
  NmsWs w = carve(workspace, S, k, max_seg_len);
  MXD_REQUIRE(workspace && workspace_bytes >= w.bytes, MXD_EWORKSPACE, "workspace %zu < %zu bytes", workspace_bytes,
              w.bytes);
  MXD_REQUIRE(((uintptr_t)workspace & 255) == 0, MXD_EINVAL, "workspace must be 256-byte aligned");
  TopkParams p = {};
  p.num_levels = 1; p.batch = S;
  p.scores[0] = dptr<float>(scores); p.seg_stride[0] = 0; p.elem_stride = 1;
  p.n[0] = max_seg_len; p.k[0] = k; p.kmax = k;
  p.seg_off = dptr<int>(seg_offsets);
  p.valid_thresh = valid_thresh;
  p.out_idx = w.order; p.out_val = nullptr; p.out_cnt = w.cnt;
  if ((rc = launch_topk(p, st, w.sortws, w.sort_bytes))) return rc;
  const bool class_aware = ids && !force_suppress;
  nms_gather_seg_kernel<<<dim3((k + 255) / 256, S), 256, 0, st>>>(dptr<float>(boxes), class_aware ? dptr<int>(ids) : nullptr,
                                                                  dptr<int>(seg_offsets), w.order, k, w.boxes,
                                                                  class_aware ? w.ids : nullptr);
  MXD_POST_LAUNCH("nms_gather_seg");
  NmsSortedArgs a = {};
  a.boxes = w.boxes; a.valid = nullptr; a.ids = class_aware ? w.ids : nullptr; a.counts = w.cnt;
  a.order = w.order; a.S = S; a.stride = k; a.n_max = k; a.thr = iou_thr; a.delta = delta;
  a.max_out = max_out; a.mask = w.mask;
  a.keep = dptr<int>(keep); a.keep_stride = cap; a.keep_cnt = dptr<int>(num_keep);
  return launch_nms_sorted(a, st);
}

struct DetWs { float4* cb; float* cs; int* cid; int* keep; void* nms; size_t bytes; };
static DetWs det_carve(void* base, long long m, int k, int cap) {
  DetWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return (char*)base + o; };
  w.cb = (float4*)take(sizeof(float4) * (size_t)m);
  w.cs = (float*)take(sizeof(float) * (size_t)m);
  w.cid = (int*)take(sizeof(int) * (size_t)m);
  w.keep = (int*)take(sizeof(int) * (size_t)(cap > 0 ? cap : 1));
  w.nms = take(carve(nullptr, 1, k > 0 ? k : 1).bytes);
  w.bytes = off;
  return w;
}
static inline void det_dims(long long n, int C, int max_per_img, long long* m, int* k, int* cap) {
  *m = n * (C - 1);
  *k = (int)std::min<long long>(*m, MXD_SORT_CAP);
  *cap = max_per_img > 0 ? std::min(*k, max_per_img) : *k;
}

size_t mxd_det_bboxes_workspace_bytes(long long n, int num_classes, int max_per_img) {
  long long m; int k, cap;
  det_dims(n, num_classes, max_per_img, &m, &k, &cap);
  return det_carve(nullptr, m, k, cap).bytes;
}

int mxd_det_bboxes(const DLTensor* boxes, const DLTensor* deltas, const DLTensor* cls_score, const float* means,
                   const float* stds, int img_h, int img_w, double wh_ratio_clip, float scale_factor, float score_thr,
                   float iou_thr, float delta, int max_per_img, DLTensor* dets, DLTensor* labels, DLTensor* num,
                   void* workspace, size_t workspace_bytes, void* stream) {
  int dev = -1, rc;
  if ((rc = check_tensor(boxes, "boxes", F32, 2, 2, &dev))) return rc;
  if ((rc = check_tensor(cls_score, "cls_score", F32, 2, 2, &dev))) return rc;
  const long long n = cls_score->shape[0];
  const int C = (int)cls_score->shape[1];
  MXD_REQUIRE(C >= 2, MXD_EINVAL, "cls_score must be (n,C) with a background column");
  const int bc = (int)boxes->shape[1];
  MXD_REQUIRE(boxes->shape[0] == n && (bc == 4 || bc == 4 * C), MXD_EINVAL, "boxes must be (n,4) or (n,4*C)");
  int dc = 0;
  if (deltas) {
    if ((rc = check_tensor(deltas, "deltas", F32, 2, 2, &dev))) return rc;
    dc = (int)deltas->shape[1];
    MXD_REQUIRE(deltas->shape[0] == n && (dc == 4 || dc == 4 * C) && bc == 4, MXD_EINVAL,
                "deltas must be (n,4) or (n,4*C) on (n,4) boxes");
    MXD_REQUIRE(means && stds && wh_ratio_clip > 0, MXD_EINVAL, "means/stds must be float[4], wh_ratio_clip > 0");
  }
  if ((rc = check_tensor(dets, "dets", F32, 2, 2, &dev))) return rc;
  if ((rc = check_tensor(labels, "labels", I32, 1, 1, &dev))) return rc;
  if ((rc = check_tensor(num, "num", I32, 1, 1, &dev))) return rc;
  long long m; int k, cap;
  det_dims(n, C, max_per_img, &m, &k, &cap);
  MXD_REQUIRE(m < (1ll << 31), MXD_ENOTSUP, "too many candidates");
  MXD_REQUIRE(dets->shape[0] == cap && dets->shape[1] == 5 && labels->shape[0] == cap && num->shape[0] >= 1, MXD_EINVAL,
              "dets / labels / num must be (%d,5) / (%d) / (1)", cap, cap);
  cudaStream_t st = as_stream(stream);
  if (m == 0 || cap == 0) {
    MXD_CUDA_OK(cudaMemsetAsync(dptr<int>(num), 0, sizeof(int), st));
    return MXD_OK;
  }
  DetWs w = det_carve(workspace, m, k, cap);
  MXD_REQUIRE(workspace && workspace_bytes >= w.bytes, MXD_EWORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, w.bytes);
  MXD_REQUIRE(((uintptr_t)workspace & 255) == 0, MXD_EINVAL, "workspace must be 256-byte aligned");
  DetF4 mu = {}, sd = {};
  if (deltas) for (int j = 0; j < 4; ++j) { mu.v[j] = means[j]; sd.v[j] = stds[j]; }
  const int clip = (img_h > 0 && img_w > 0) ? 1 : 0;
  det_candidates_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(
      dptr<float>(boxes), bc, deltas ? dptr<float>(deltas) : nullptr, dc, dptr<float>(cls_score), (int)n, C, mu, sd,
      deltas ? (float)fabs(log(wh_ratio_clip)) : 0.0f, (float)(img_h - 1), (float)(img_w - 1), clip, scale_factor, w.cb, w.cs, w.cid);
  MXD_POST_LAUNCH("det_candidates");
  NmsWs nw = carve(w.nms, 1, k);
  TopkParams p = {};
  p.num_levels = 1; p.batch = 1;
  p.scores[0] = w.cs; p.seg_stride[0] = m; p.elem_stride = 1;
  p.n[0] = (int)m; p.k[0] = k; p.kmax = k;
  p.valid_thresh = score_thr;
  p.out_idx = nw.order; p.out_cnt = nw.cnt;
  if ((rc = launch_topk(p, st))) return rc;
  nms_gather_kernel<<<(k + 255) / 256, 256, 0, st>>>(reinterpret_cast<const float*>(w.cb), w.cid, nw.order, k, nw.boxes, nw.ids);
  MXD_POST_LAUNCH("nms_gather");
  NmsSortedArgs a = {};
  a.boxes = nw.boxes; a.ids = nw.ids; a.counts = nw.cnt; a.order = nw.order;
  a.S = 1; a.stride = k; a.n_max = k; a.thr = iou_thr; a.delta = delta; a.max_out = cap;
  a.mask = nw.mask; a.keep = w.keep; a.keep_stride = cap; a.keep_cnt = dptr<int>(num);
  if ((rc = launch_nms_sorted(a, st))) return rc;
  det_output_kernel<<<(cap + 255) / 256, 256, 0, st>>>(w.cb, w.cs, w.cid, w.keep, cap, dptr<float>(dets), dptr<int>(labels));
  MXD_POST_LAUNCH("det_output");
  return MXD_OK;
}

size_t mxd_box_nms_workspace_bytes(int batch, int n, int topk) { return carve(nullptr, batch, eff_k(n, topk), n).bytes; }

int mxd_box_nms(const DLTensor* data, DLTensor* out, DLTensor* index, float overlap_thresh, float valid_thresh,
                int topk, int coord_start, int score_index, int id_index, int force_suppress, int in_format,
                int out_format, void* workspace, size_t workspace_bytes, void* stream) {
  int dev = -1, rc;
  if ((rc = check_tensor(data, "data", F32, 2, 3, &dev))) return rc;
  if ((rc = check_tensor(out, "out", F32, data->ndim, data->ndim, &dev))) return rc;
  const int B = data->ndim == 3 ? (int)data->shape[0] : 1;
  const long long N = data->shape[data->ndim - 2];
  const int K = (int)data->shape[data->ndim - 1];
  MXD_REQUIRE(numel(out) == numel(data), MXD_EINVAL, "out must have the shape of data");
  MXD_REQUIRE(K >= 5 && coord_start >= 0 && coord_start + 4 <= K && score_index >= 0 && score_index < K &&
              id_index < K, MXD_EINVAL, "bad coord_start/score_index/id_index for K=%d", K);
  MXD_REQUIRE((in_format == 0 || in_format == 1) && (out_format == 0 || out_format == 1), MXD_EINVAL, "bad format");
  if (index) {
    if ((rc = check_tensor(index, "index", I32, data->ndim - 1, data->ndim - 1, &dev))) return rc;
    MXD_REQUIRE(numel(index) == (int64_t)B * N, MXD_EINVAL, "index must be (B,N)");
  }
  if (B == 0 || N == 0) return MXD_OK;
  const int k = eff_k(N, topk);
  NmsWs w = carve(workspace, B, k, N);
  MXD_REQUIRE(workspace && workspace_bytes >= w.bytes, MXD_EWORKSPACE, "workspace %zu < %zu bytes",
              workspace_bytes, w.bytes);
  cudaStream_t st = as_stream(stream);
  TopkParams p = {};
  p.num_levels = 1; p.batch = B;
  p.scores[0] = dptr<float>(data) + score_index; p.seg_stride[0] = N * K; p.elem_stride = K;
  p.n[0] = (int)N; p.k[0] = k; p.kmax = k;
  p.valid_thresh = valid_thresh;
  p.out_idx = w.order; p.out_cnt = w.cnt;
  if ((rc = launch_topk(p, st, w.sortws, w.sort_bytes))) return rc;
  const bool class_aware = id_index >= 0 && !force_suppress;
  dim3 g1((k + 255) / 256, B);
  boxnms_gather_kernel<<<g1, 256, 0, st>>>(dptr<float>(data), w.order, (int)N, K, k, coord_start,
                                           class_aware ? id_index : -1, in_format, w.boxes,
                                           class_aware ? w.ids : nullptr);
  MXD_POST_LAUNCH("boxnms_gather");
  NmsSortedArgs a = {};
  a.boxes = w.boxes; a.ids = class_aware ? w.ids : nullptr; a.counts = w.cnt; a.order = w.order;
  a.S = B; a.stride = k; a.n_max = k; a.thr = overlap_thresh; a.delta = 0.0f; a.max_out = -1;
  a.mask = w.mask; a.keep = w.keep; a.keep_stride = k; a.keep_cnt = w.keep_cnt;
  if ((rc = launch_nms_sorted(a, st))) return rc;
  dim3 g2(((int)N + 255) / 256, B);
  boxnms_write_kernel<<<g2, 256, 0, st>>>(dptr<float>(data), w.keep, w.keep_cnt, (int)N, K, k, coord_start,
                                          in_format, out_format, dptr<float>(out),
                                          index ? dptr<int>(index) : nullptr);
  MXD_POST_LAUNCH("boxnms_write");
  return MXD_OK;
}

int mxd_box_nms_backward(const DLTensor* out_grad, const DLTensor* index, DLTensor* in_grad, void* stream) {
  int dev = -1, rc;
  if ((rc = check_tensor(out_grad, "out_grad", F32, 2, 3, &dev))) return rc;
  if ((rc = check_tensor(in_grad, "in_grad", F32, out_grad->ndim, out_grad->ndim, &dev))) return rc;
  if ((rc = check_tensor(index, "index", I32, out_grad->ndim - 1, out_grad->ndim - 1, &dev))) return rc;
  const int B = out_grad->ndim == 3 ? (int)out_grad->shape[0] : 1;
  const int N = (int)out_grad->shape[out_grad->ndim - 2], K = (int)out_grad->shape[out_grad->ndim - 1];
  MXD_REQUIRE(numel(in_grad) == numel(out_grad) && numel(index) == (int64_t)B * N, MXD_EINVAL, "shape mismatch");
  if (B == 0 || N == 0) return MXD_OK;
  cudaStream_t st = as_stream(stream);
  MXD_CUDA_OK(cudaMemsetAsync(dptr<float>(in_grad), 0, sizeof(float) * (size_t)numel(in_grad), st));
  dim3 g((N + 255) / 256, B);
  boxnms_backward_kernel<<<g, 256, 0, st>>>(dptr<float>(out_grad), dptr<int>(index), N, K, dptr<float>(in_grad));
  MXD_POST_LAUNCH("boxnms_backward");
  return MXD_OK;
}

}  // extern "C"
