// Tile-resident RoIAlign backward for B200 (Spec A backward, Spec G).
//
// Contract: _backward_ROIAlign of mxnet 1.3.0 as used by mxdetection/ops
// (/root/reference/README.md:24) and SingleLevelRoI (/root/reference/README.md:32).
//
// Why: global RED.ADD is bound at ~200 G sector-ops/s on B200 (profiles/microbench), which
// caps an atomic scatter at ~0.5 ms on BASELINE config 3; shared-memory fp32 atomics are CAS
// loops.  This kernel needs neither:
//
//  * every gradient map is cut into tiles of <= 25 (7x7) / 19 (14x14) rows x 48 columns x 32 channels that live in
//    shared memory as [pixel][33 words] (channel innermost).  Lane = channel, so all 32 lanes of
//    a warp run the same control flow with warp-uniform weights, and every shared-memory access
//    is bank-conflict free;
//  * warp w of the CTA OWNS tile row w (fixed row pitch, whatever the level): only that warp ever
//    touches it, so the accumulation is a plain LDS / FFMA / STS - no atomics and no barrier at all;
//  * the planner (one warp per RoI) turns Spec A's 2 x PH*sr row taps into a dense per-row
//    table (first bin, up to 7 bin weights), lists the RoIs intersecting every tile, and packs
//    the x taps; a producer warp streams, per (RoI, tile) pair, the RoI's 32-channel slice of
//    grad_out (one TMA bulk copy, read as g[lane*bins + bin]: conflict free for 7x7 because bins is
//    odd; 14x14 reads bin pairs), the row-table slice and the tile-relative x taps through an mbarrier ring;
//  * each tile is written to HBM exactly once with plain coalesced stores (req=write needs no
//    memset; req=add adds on the way out).
//
// RoIs that sample outside the image, are taller than 256 feature rows or span more than 64 tiles
// take the generic RED path afterwards (gather_roi_chunk<true>) - they are rare.
#include <stdlib.h>
#include "roi_align.cuh"
#include "ptx.cuh"

namespace mxd {

#ifndef MXD_TB_R
#define MXD_TB_R 1
#endif
#ifndef MXD_TB_CTAS          // CTAs per SM (1: one 25-row tile + 9 stages; 2: two independent 12-row pipelines)
#define MXD_TB_CTAS 1
#endif
#ifndef MXD_TB_ROWS
#define MXD_TB_ROWS (MXD_TB_CTAS == 1 ? 25 : 12)
#endif
#ifndef MXD_TB_STAGES
#define MXD_TB_STAGES (MXD_TB_CTAS == 1 ? 9 : 4)
#endif
#ifndef MXD_TB_PRODUCERS
#define MXD_TB_PRODUCERS (MXD_TB_CTAS == 1 ? 3 : 2)       /* 25 + 3 warps = 896 threads: 72 registers per thread */
#endif
constexpr int kTbR = MXD_TB_R;                     // tile rows per consumer warp (1; 2 measured 5 % slower: fewer, fatter warps)
constexpr int kTbWarps = (MXD_TB_ROWS + kTbR - 1) / kTbR;   // consumer warps; warp w owns tile rows R*w .. R*w+R-1 (a tile has <= 25 rows)
constexpr int kTbProducers = MXD_TB_PRODUCERS;                    // producer warps: message m is built by producer m mod 4 (one warp needs
                                                   // ~1500 cycles per message - two bulk copies, a barrier wait, the packed
                                                   // bins - and left the consumers waiting 30 % of the time)
constexpr int kTbThreads = (kTbWarps + kTbProducers) * 32;
constexpr int kTbMaxStages = 16;                   // ring depth is per configuration (9 for 7x7, 4 for 14x14)
constexpr int kTbMaxHfCap = 256;                   // rows of the dense per-RoI row table: min(cap, tallest map)
constexpr int kTbPix = 33;                         // words per tile pixel (32 channels + 1 pad)
constexpr int kTbMaxTiles = 64;                    // tiles one RoI may intersect
constexpr int kTbMaxTh = kTbR * kTbWarps;
constexpr int kTbMaxTw = 48;
constexpr int kTbRowWords = (kTbMaxTw + 1) * kTbPix;   // fixed row pitch: 48 columns + the trash column
constexpr int kTbTrash = kTbMaxTw * kTbPix * 4;        // byte offset of the trash pixel inside a row
constexpr int kTbSmem = MXD_TB_CTAS == 1 ? 227 * 1024 : 113 * 1024;
constexpr int kTbCtlBytes = 384;                   // TbCtl

enum { kMsgPair = 0, kMsgBegin = 1, kMsgZero = 2, kMsgStop = 3 };

struct TLevel { int H, W, th, tw, nty, ntx, tile_base; };

struct TCfg {
  TLevel lv[MXD_MAX_LEVELS];
  int L, N, C, PH, PW, sr, ty, tx, bins;
  int tiles_per_img, NT, ncg, n_items;
  int stage_bytes, off_xt, off_rt, off_g, tile_bytes, smem_bytes, max_rows, max_hf, n_stages;
  int accumulate;
  int level_mask;      // bit l set: level l is this kernel's (the other levels are neither zero-filled nor accumulated here)
  float finest, inv_count;
};

struct TWs {
  int* hdr;        // [0] item counter, [1] fallback count
  int* cnt;        // [NT] RoIs per tile
  int* cursor;     // [NT]
  int* start;      // [NT]
  int* fb_list;    // [R]
  int4* roihdr;    // [R][2]  {y0, Hf, x0, x1} {b, lvl, ok, ntiles}
  uint2* xtab;     // [R][tx] {column of the low tap (unclamped form), weight of the high tap}
  uint4* rowtab;   // [R][max_hf][2]  {first bin | nbins<<8, w0..w6}
  int* pairs;      // [R * kTbMaxTiles]  per-tile RoI lists, ascending RoI index (deterministic summation order)
  int* pairs_raw;  // [R * kTbMaxTiles]  the same lists in the order the grouping atomics happened to fill them
  size_t bytes;
};

static TWs carve_tile(void* base, int R, int NT, int tx, int max_hf) {
  TWs w;
  size_t off = 0;
  const size_t r1 = (size_t)(R > 0 ? R : 1);
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return (char*)base + o; };
  w.hdr = (int*)take(sizeof(int) * 64);
  w.cnt = (int*)take(sizeof(int) * (size_t)NT);
  w.cursor = (int*)take(sizeof(int) * (size_t)NT);
  w.start = (int*)take(sizeof(int) * (size_t)NT);
  w.fb_list = (int*)take(sizeof(int) * r1);
  w.roihdr = (int4*)take(sizeof(int4) * 2 * r1);
  w.xtab = (uint2*)take(sizeof(uint2) * r1 * tx);
  w.rowtab = (uint4*)take(sizeof(uint4) * 2 * r1 * max_hf);
  w.pairs = (int*)take(sizeof(int) * r1 * kTbMaxTiles);
  w.pairs_raw = (int*)take(sizeof(int) * r1 * kTbMaxTiles);
  w.bytes = off;
  return w;
}

// Host: tile geometry per level for the shared-memory budget.  7x7 and 14x14 at sample_ratio 2 with
// C % 32 == 0 are instantiated; everything else keeps the RED kernels.
static bool make_tcfg(int N, int C, int L, const int* Hs, const int* Ws, int PH, int PW, int sr, float finest,
                      int accumulate, TCfg* c) {
  if (sr != 2 || !((PH == 7 && PW == 7) || (PH == 14 && PW == 14)) || (C & 31) != 0 || C == 0) return false;
  memset(c, 0, sizeof(*c));
  c->L = L; c->N = N; c->C = C; c->PH = PH; c->PW = PW; c->sr = sr; c->ty = PH * sr; c->tx = PW * sr;
  c->bins = PH * PW; c->finest = finest; c->inv_count = 1.0f / (float)(sr * sr); c->accumulate = accumulate;
  c->level_mask = ~0;
  c->off_xt = 32;                                   // PW x {w1, w2, w3, 4 x u8 column}
  c->off_rt = (int)align_up((size_t)c->off_xt + PW * 16, 32);
  c->off_g = c->off_rt + kTbMaxTh * 32;
  c->stage_bytes = (int)align_up((size_t)c->off_g + 32 * c->bins * 4, 128);
  c->n_stages = PW == 7 ? MXD_TB_STAGES : (MXD_TB_CTAS == 1 ? 4 : 2);   // 7.5 KB / 26 KB of grad_out per stage
  c->tile_bytes = (kTbSmem - c->n_stages * c->stage_bytes - kTbCtlBytes) & ~15;
  const int max_rows = c->tile_bytes / (kTbRowWords * 4);
  c->max_rows = max_rows;
  int base = 0;
  for (int l = 0; l < L; ++l) {
    TLevel& v = c->lv[l];
    v.H = Hs[l]; v.W = Ws[l];
    if (v.H > c->max_hf) c->max_hf = v.H < kTbMaxHfCap ? v.H : kTbMaxHfCap;
    v.ntx = (v.W + kTbMaxTw - 1) / kTbMaxTw;
    v.tw = (v.W + v.ntx - 1) / v.ntx;
    int th_cap = max_rows;
    if (th_cap > kTbMaxTh) th_cap = kTbMaxTh;
    if (th_cap < 1) return false;
    v.nty = (v.H + th_cap - 1) / th_cap;
    v.th = (v.H + v.nty - 1) / v.nty;
    v.tile_base = base;
    base += v.nty * v.ntx;
  }
  c->tiles_per_img = base;
  c->NT = N * base;
  c->ncg = C / 32;
  if ((long long)c->NT * c->ncg > 0x3fffffffLL) return false;
  c->n_items = c->NT * c->ncg;
  c->smem_bytes = c->tile_bytes + c->n_stages * c->stage_bytes + kTbCtlBytes;
  return c->NT > 0;
}

// ------------------------------------------------------------------- planner ------
// One warp per RoI: Spec A sample tables -> dense row table, packed x taps, tile counts.
template <int PH>
__global__ void __launch_bounds__(256) tplan_rois_kernel(FpnDesc d, TCfg c, TWs w, const float* __restrict__ rois,
                                                          const int* __restrict__ levels, int R) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  pdl_wait();                       // the workspace memset (and whatever produced rois) is complete
  pdl_launch_dependents();
  if (n >= R) return;
  const unsigned full = 0xffffffffu;
  const RoiGeom g = roi_geom(d, rois, levels, n, c.PH, c.PW, c.sr, c.finest);
  bool ok = g.ok && g.H >= 2 && g.W >= 2;
  if (g.ok && !((c.level_mask >> g.lvl) & 1)) {      // another kernel owns this level
    if (lane == 0) w.roihdr[(size_t)n * 2 + 1] = make_int4(0, 0, 0, 0);
    return;
  }
  int ylo = 0x3fffffff, xlo = 0x3fffffff;
  float yl = 0.0f, xl = 0.0f;
  bool valid = true;
  if (lane < c.ty) {
    const AxisTap a = axis_tap(g.rsh, g.bh, c.sr, lane / c.sr, lane % c.sr, g.H, 1);
    valid = valid && a.valid;
    const bool cl = a.hi == a.lo;      // clamped at the border -> unclamped form (size-2, l = 1)
    ylo = cl ? a.lo - 1 : a.lo;
    yl = cl ? 1.0f : a.l;
  }
  if (lane < c.tx) {
    const AxisTap a = axis_tap(g.rsw, g.bw, c.sr, lane / c.sr, lane % c.sr, g.W, 1);
    valid = valid && a.valid;
    const bool cl = a.hi == a.lo;
    xlo = cl ? a.lo - 1 : a.lo;
    xl = cl ? 1.0f : a.l;
  }
  ok = ok && __all_sync(full, valid);
  const int y0 = __reduce_min_sync(full, ylo);
  const int y1 = __reduce_max_sync(full, lane < c.ty ? ylo + 1 : -1);
  const int x0 = __reduce_min_sync(full, xlo);
  const int x1 = __reduce_max_sync(full, lane < c.tx ? xlo + 1 : -1);
  const int Hf = y1 - y0 + 1;
  ok = ok && Hf >= 1 && Hf <= c.max_hf && y0 >= 0 && x0 >= 0;
  bool rows_ok = true;
  if (ok) {
    // dense rows: bin weights accumulate in PH registers (static index: sr = 2 here), then shift to the first
    // non-zero bin - no local-memory array.  A row fed by more than 7 bins (bin height < 0.3 px) goes to the fallback.
    for (int i0 = 0; i0 < Hf; i0 += 32) {
      const int i = i0 + lane;
      const int row = y0 + i;
      float wabs[PH];
#pragma unroll
      for (int k = 0; k < PH; ++k) wabs[k] = 0.0f;
#pragma unroll
      for (int t = 0; t < 2 * PH; ++t) {
        const int lo_t = __shfl_sync(full, ylo, t);
        const float l_t = __shfl_sync(full, yl, t);
        float wt = 0.0f;
        if (lo_t == row) wt = 1.0f - l_t;
        else if (lo_t + 1 == row) wt = l_t;
        wabs[t >> 1] += wt;
      }
      int pa = -1, pl = -1;
#pragma unroll
      for (int k = 0; k < PH; ++k)
        if (wabs[k] != 0.0f) { if (pa < 0) pa = k; pl = k; }
      if (i < Hf) {
        const int nph = pa < 0 ? 0 : pl - pa + 1;
        const int p0 = pa < 0 ? 0 : pa;
        if (nph > 7) rows_ok = false;
        float wr[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) {
          float v = 0.0f;
#pragma unroll
          for (int j = 0; j < PH; ++j) if (j == p0 + k) v = wabs[j];
          wr[k] = v;
        }
        uint4* rec = w.rowtab + ((size_t)n * c.max_hf + i) * 2;
        rec[0] = make_uint4((unsigned)p0 | ((unsigned)nph << 8), __float_as_uint(wr[0]),
                            __float_as_uint(wr[1]), __float_as_uint(wr[2]));
        rec[1] = make_uint4(__float_as_uint(wr[3]), __float_as_uint(wr[4]), __float_as_uint(wr[5]),
                            __float_as_uint(wr[6]));
      }
    }
  }
  ok = ok && __all_sync(full, rows_ok);
  int tya = 0, tyb = 0, txa = 0, txb = 0, nt = 0;
  if (ok) {
    const TLevel& v = c.lv[g.lvl];
    tya = y0 / v.th; tyb = y1 / v.th; txa = x0 / v.tw; txb = x1 / v.tw;
    nt = (tyb - tya + 1) * (txb - txa + 1);
    if (nt > kTbMaxTiles) ok = false;
  }
  if (!ok) {
    if (lane == 0) {
      w.roihdr[(size_t)n * 2 + 1] = make_int4(0, 0, 0, 0);
      w.fb_list[atomicAdd(&w.hdr[1], 1)] = n;
    }
    return;
  }
  if (lane < c.tx) w.xtab[(size_t)n * c.tx + lane] = make_uint2((unsigned)xlo, __float_as_uint(xl));
  const TLevel& v = c.lv[g.lvl];
  const int ncol = txb - txa + 1;
  for (int q = lane; q < nt; q += 32) {
    const int t = g.b * c.tiles_per_img + v.tile_base + (tya + q / ncol) * v.ntx + (txa + q % ncol);
    atomicAdd(&w.cnt[t], 1);
  }
  if (lane == 0) {
    w.roihdr[(size_t)n * 2] = make_int4(y0, Hf, x0, x1);
    w.roihdr[(size_t)n * 2 + 1] = make_int4(g.b, g.lvl, 1, nt);
  }
}

// Exclusive scan of the tile counts and the per-tile RoI lists (single CTA).  The list cursors live in shared
// memory when the tile table fits (global atomics put two L2 round trips on every RoI of the fill loop).
constexpr int kTbSmemTiles = 4096;

__global__ void __launch_bounds__(1024) tplan_group_kernel(TCfg c, TWs w, int R) {
  __shared__ int s_part[1024];
  __shared__ int s_pos[kTbSmemTiles];      // next free slot of every tile's list
  const int tid = threadIdx.x;
  pdl_wait();                       // the RoI plans are complete
  pdl_launch_dependents();
  const bool in_smem = c.NT <= kTbSmemTiles;
  const int per = (c.NT + 1023) / 1024;
  int sum = 0;
  for (int i = tid * per; i < min(c.NT, (tid + 1) * per); ++i) sum += w.cnt[i];
  s_part[tid] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int v = tid >= o ? s_part[tid - o] : 0;
    __syncthreads();
    s_part[tid] += v;
    __syncthreads();
  }
  int run = s_part[tid] - sum;
  for (int i = tid * per; i < min(c.NT, (tid + 1) * per); ++i) {
    w.start[i] = run;
    if (in_smem) s_pos[i] = run;
    run += w.cnt[i];
  }
  __syncthreads();
  __threadfence_block();
  for (int n = tid; n < R; n += 1024) {
    const int4 h1 = w.roihdr[(size_t)n * 2 + 1];
    if (!h1.z) continue;
    const int4 h0 = w.roihdr[(size_t)n * 2];
    const TLevel& v = c.lv[h1.y];
    const int tya = h0.x / v.th, tyb = (h0.x + h0.y - 1) / v.th, txa = h0.z / v.tw, txb = h0.w / v.tw;
    for (int ty = tya; ty <= tyb; ++ty)
      for (int tx = txa; tx <= txb; ++tx) {
        const int t = h1.x * c.tiles_per_img + v.tile_base + ty * v.ntx + tx;
        const int pos = in_smem ? atomicAdd(&s_pos[t], 1) : w.start[t] + atomicAdd(&w.cursor[t], 1);
        w.pairs_raw[pos] = n;
      }
  }
}

// Orders every tile's list by RoI index: the consumers add the RoIs of a tile in list order, so a list in atomic-
// arrival order makes the gradient differ in the last bits from run to run.  One warp per tile: the ids are marked in
// a 4096-bit shared-memory bitmap (window by window over the list's id range - one window when a tile's RoIs are
// neighbours in the RoI array, as the RoIs of one image usually are) and read back in ascending order.
__global__ void __launch_bounds__(256) tplan_sort_kernel(TCfg c, TWs w) {
  __shared__ unsigned s_bm[8][128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_wait();                       // the raw lists are complete
  pdl_launch_dependents();
  unsigned* bm = s_bm[warp];
  const int t = blockIdx.x * 8 + warp;
  if (t >= c.NT) return;
  const int cnt = w.cnt[t];
  if (cnt == 0) return;
  const int st = w.start[t];
  int lo = 0x7fffffff, hi = -1;
  for (int i = lane; i < cnt; i += 32) { const int id = w.pairs_raw[st + i]; lo = min(lo, id); hi = max(hi, id); }
  lo = __reduce_min_sync(0xffffffffu, lo);
  hi = __reduce_max_sync(0xffffffffu, hi);
  int wpos = st;
  for (int base = lo; base <= hi; base += 4096) {
#pragma unroll
    for (int k = 0; k < 4; ++k) bm[lane * 4 + k] = 0u;
    __syncwarp();
    for (int i = lane; i < cnt; i += 32) {
      const unsigned r = (unsigned)(w.pairs_raw[st + i] - base);
      if (r < 4096u) atomicOr(&bm[r >> 5], 1u << (r & 31));
    }
    __syncwarp();
    unsigned wd[4];
    int mine = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { wd[k] = bm[lane * 4 + k]; mine += __popc(wd[k]); }
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    int off = wpos + incl - mine;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      unsigned bits = wd[k];
      while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1;
        w.pairs[off++] = base + (lane * 4 + k) * 32 + b;
      }
    }
    wpos += __shfl_sync(0xffffffffu, incl, 31);
    __syncwarp();
  }
}

// --------------------------------------------------------------- main kernel ------
struct TbCtl {
  u64 full[kTbMaxStages];
  u64 empty[kTbMaxStages];
  int items[2];          // item ids, written by producer 0 and read by the others (double-buffered)
};
static_assert(sizeof(TbCtl) <= kTbCtlBytes, "control block");

// The 4 x taps of bin pw (2 samples x lo/hi) as up to 4 DISTINCT tile columns with merged weights
// (already scaled by 1/count), so the consumer can issue the 4 loads before the 4 stores.  Columns
// outside the tile, and unused slots, are the row's trash column (index kTbMaxTw).  Packed to 16 bytes
// (a uniform LDS.128 costs 2 shared-memory wavefronts, the scarce resource of this kernel):
// {w1, w2, w3, c0 | c1<<8 | c2<<16 | c3<<24}; w0 = 2/count - (w1 + w2 + w3).
__device__ __forceinline__ uint4 tb_pack_bin(uint4 e, int tx0, int tw, float inv_count) {
  const int a = (int)e.x - tx0, b = (int)e.z - tx0;       // low-tap columns of samples A, B (b >= a)
  const float lA = __uint_as_float(e.y) * inv_count, lB = __uint_as_float(e.w) * inv_count;
  const float hB = inv_count - lB;
  int c0 = a, c1 = a + 1, c2, c3;
  float w1 = lA, w2, w3;
  if (b == a) { w1 += lB; c2 = -1; c3 = -1; w2 = 0.0f; w3 = 0.0f; }
  else if (b == a + 1) { w1 += hB; c2 = b + 1; w2 = lB; c3 = -1; w3 = 0.0f; }
  else { c2 = b; c3 = b + 1; w2 = hB; w3 = lB; }
  auto col = [&](int cx) { return ((unsigned)cx < (unsigned)tw) ? (unsigned)cx : (unsigned)kTbMaxTw; };
  return make_uint4(__float_as_uint(w1), __float_as_uint(w2), __float_as_uint(w3),
                    col(c0) | (col(c1) << 8) | (col(c2) << 16) | (col(c3) << 24));
}

// Producer p of kTbProducers.  All producers walk the same sequence of items (producer 0 pulls the item id off the
// global counter and passes it on through shared memory, one named barrier per item) and therefore number the messages
// alike; producer p builds the messages m = p mod kTbProducers into stage m mod n_stages.
__device__ __forceinline__ void tb_producer(const FpnDesc& d, const TCfg& c, const TWs& w,
                                            const float* __restrict__ gout, unsigned char* stages, TbCtl* ctl,
                                            int lane, int p) {
  const unsigned full = 0xffffffffu;
  constexpr int P = kTbProducers;
  const int S = c.n_stages;
  int m = 0;            // messages issued so far by all producers
  int s = p % S;        // stage of this producer's next message (message number == p mod P)
  uint32_t par = 1;     // parity to wait for on empty[s]: the first pass over the ring succeeds at once
  int mine = p;         // number of this producer's next message
  auto advance = [&]() { mine += P; s += P; if (s >= S) { s -= S; par ^= 1u; } };
  const uint32_t g_bytes = (uint32_t)(32 * c.bins * 4);
#ifdef MXD_TB_PROF
  long long t_emp = 0, t_all = clock64();
#define TB_WAIT_EMPTY() { const long long _t = clock64(); mbar_wait(&ctl->empty[s], par); t_emp += clock64() - _t; }
#else
#define TB_WAIT_EMPTY() mbar_wait(&ctl->empty[s], par)
#endif
  for (int it = 0;; ++it) {
    if (p == 0 && lane == 0) ctl->items[it & 1] = atomicAdd(&w.hdr[0], 1);
    asm volatile("bar.sync 2, %0;" ::"n"(32 * kTbProducers) : "memory");
    const int item = ctl->items[it & 1];
    if (item >= c.n_items) break;
    // Items are level-major, coarsest level first (its tiles carry the most RoIs: heavy items early, light ones
    // fill the tail); inside a level all tiles of one (image, channel group) are neighbours, so the RoIs'
    // grad_out slices - every RoI lives on exactly one level - are re-read from L2 while they are hot.
    int l = c.L - 1, rem = item;
    for (;;) {
      const int n_l = c.lv[l].nty * c.lv[l].ntx * c.N * c.ncg;
      if (rem < n_l || l == 0) break;
      rem -= n_l; --l;
    }
    if (!((c.level_mask >> l) & 1)) continue;     // (same decision in every producer)
    const TLevel& v = c.lv[l];
    const int tiles_l = v.nty * v.ntx;
    const int tl = rem % tiles_l;
    const int r = rem / tiles_l;
    const int cg = r % c.ncg, b = r / c.ncg;
    const int t = v.tile_base + tl;
    const int ty0 = (tl / v.ntx) * v.th, tx0 = (tl % v.ntx) * v.tw;
    const int tile_id = b * c.tiles_per_img + t;
    const int cnt = w.cnt[tile_id], start = w.start[tile_id];
    if (cnt == 0 && c.accumulate) continue;    // req=add and nothing to add
    if (mine == m) {                           // the item's first message: begin (RoIs follow) or zero (write zeros)
      TB_WAIT_EMPTY();
      if (lane == 0) {
        int* hd = reinterpret_cast<int*>(stages + (size_t)s * c.stage_bytes);
        hd[0] = cnt ? kMsgBegin : kMsgZero;
        reinterpret_cast<int4*>(hd)[1] = make_int4(l, b, cg * 32, ty0 | (tx0 << 16));
        mbar_arrive(&ctl->full[s]);
      }
      advance();
    }
    ++m;
    // pairs j = j0, j0 + P, ... of this item are mine; 32 of them are looked up at a time (lane i: pair j0 + i * P)
    for (int j0 = mine - m; j0 < cnt; j0 += 32 * P) {
      const int jl = j0 + lane * P;
      const int my_n = jl < cnt ? w.pairs[start + jl] : 0;
      const int4 h0 = w.roihdr[(size_t)my_n * 2];
      if (jl < cnt)   // the slices of my next pairs start their trip HBM -> L2 now
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gout + ((size_t)my_n * c.C + cg * 32) * c.bins),
                     "r"(g_bytes)
                     : "memory");
      const int nb = min(32, (cnt - j0 + P - 1) / P);
      uint4 xe = make_uint4(0u, 0u, 0u, 0u);      // lane pw: samples 2pw, 2pw+1 of the next pair's RoI
      {
        const int n0 = __shfl_sync(full, my_n, 0);
        if (lane < c.PW) xe = reinterpret_cast<const uint4*>(w.xtab + (size_t)n0 * c.tx)[lane];
      }
      for (int q = 0; q < nb; ++q) {
        const int n = __shfl_sync(full, my_n, q);
        const int y0 = __shfl_sync(full, h0.x, q), Hf = __shfl_sync(full, h0.y, q);
        const uint4 xc = xe;
        if (q + 1 < nb) {
          const int n1 = __shfl_sync(full, my_n, q + 1);
          if (lane < c.PW) xe = reinterpret_cast<const uint4*>(w.xtab + (size_t)n1 * c.tx)[lane];
        }
        TB_WAIT_EMPTY();
        unsigned char* st = stages + (size_t)s * c.stage_bytes;
        const int ra = max(y0, ty0) - ty0;
        const int rb = min(y0 + Hf - 1, ty0 + v.th - 1) - ty0;
        if (lane < c.PW) reinterpret_cast<uint4*>(st + c.off_xt)[lane] = tb_pack_bin(xc, tx0, v.tw, c.inv_count);
        const uint32_t rt_bytes = (uint32_t)((rb - ra + 1) * 32);
        if (lane == 0) {
          reinterpret_cast<int*>(st)[0] = kMsgPair | (ra << 8) | (rb << 16);
          mbar_arrive_expect_tx(&ctl->full[s], g_bytes + rt_bytes);
        }
        __syncwarp();
        // lane 0 copies the grad_out slice, lane 1 the row-table slice: one instruction issues both
        if (lane == 0)
          bulk_g2s(st + c.off_g, gout + ((size_t)n * c.C + cg * 32) * c.bins, g_bytes, &ctl->full[s]);
        else if (lane == 1)
          bulk_g2s(st + c.off_rt + ra * 32, w.rowtab + ((size_t)n * c.max_hf + (ra + ty0 - y0)) * 2, rt_bytes,
                   &ctl->full[s]);
        advance();
      }
    }
    m += cnt;
  }
  if (mine == m) {
    TB_WAIT_EMPTY();
    if (lane == 0) {
      reinterpret_cast<int*>(stages + (size_t)s * c.stage_bytes)[0] = kMsgStop;
      mbar_arrive(&ctl->full[s]);
    }
  }
#ifdef MXD_TB_PROF
  if (lane == 0 && p == 0) {
    atomicAdd(reinterpret_cast<unsigned long long*>(w.hdr + 16), (unsigned long long)t_emp);
    atomicAdd(reinterpret_cast<unsigned long long*>(w.hdr + 18), (unsigned long long)(clock64() - t_all));
  }
#endif
}

// One tile row of one RoI for this lane's channel, part 1: h[pw] = sum_ph Wy[row][ph] * g[ph][pw].  Returns false
// when no bin feeds the row.
template <int PW>
__device__ __forceinline__ bool tb_h(const float* __restrict__ gl, const uint2* __restrict__ rt, float (&h)[PW]) {
  const uint2 q0 = rt[0];                    // {first bin | nbins << 8, w0}
  const int nph = (int)((q0.x >> 8) & 0xffu);
  if (nph == 0) return false;
  const float* gp = gl + (int)(q0.x & 0xffu) * PW;
  // g[lane * bins + bin]: conflict-free 32-bit loads when bins is odd (7x7); for 14x14 (bins = 196, lane stride
  // = 4 banks) 64-bit loads of bin pairs halve the 4-way conflicts (rows of PW = 14 floats are 8-byte aligned)
  auto grow = [&](int k, float (&gv)[PW]) {
    if constexpr ((PW & 1) == 0) {
      const float2* g2 = reinterpret_cast<const float2*>(gp + k * PW);
#pragma unroll
      for (int j = 0; j < PW / 2; ++j) { const float2 v = g2[j]; gv[2 * j] = v.x; gv[2 * j + 1] = v.y; }
    } else {
#pragma unroll
      for (int pw = 0; pw < PW; ++pw) gv[pw] = gp[k * PW + pw];
    }
  };
  {
    const float w0 = __uint_as_float(q0.y);
    float gv[PW];
    grow(0, gv);
#pragma unroll
    for (int pw = 0; pw < PW; ++pw) h[pw] = w0 * gv[pw];
  }
  if (nph > 1) {
    const uint2 q1 = rt[1];                  // {w1, w2}
    const float w1 = __uint_as_float(q1.x);
    {
      float gv[PW];
      grow(1, gv);
#pragma unroll
      for (int pw = 0; pw < PW; ++pw) h[pw] = fmaf(w1, gv[pw], h[pw]);
    }
    if (nph > 2) {
      const float w2 = __uint_as_float(q1.y);
      {
        float gv[PW];
        grow(2, gv);
#pragma unroll
        for (int pw = 0; pw < PW; ++pw) h[pw] = fmaf(w2, gv[pw], h[pw]);
      }
      if (nph > 3) {
        const uint4 q2 = reinterpret_cast<const uint4*>(rt)[1];   // {w3..w6}
        const float wk[4] = {__uint_as_float(q2.x), __uint_as_float(q2.y), __uint_as_float(q2.z), __uint_as_float(q2.w)};
#pragma unroll
        for (int k = 3; k < 7; ++k) {
          if (k < nph) {
            float gv[PW];
            grow(k, gv);
#pragma unroll
            for (int pw = 0; pw < PW; ++pw) h[pw] = fmaf(wk[k - 3], gv[pw], h[pw]);
          }
        }
      }
    }
  }
  return true;
}

// Part 2, for the warp's two rows at once: the (up to) 4 distinct columns of every bin get w_k * h[pw] with plain
// read-modify-writes - the rows belong to this warp, and distinct columns let the loads issue before the stores.
// The packed bin {w1, w2, w3, 4 column bytes} is fetched and unpacked once for both rows.  (Skipping the slots that
// point at the trash column - 19 % of them on the benchmark workload - with warp-uniform branches measured 5 %
// SLOWER: 0.437 vs 0.414 ms; the straight-line bin body schedules better than it saves.  Predicating them off in
// inline PTX (@p ld / fma / st.shared, no branch) changed nothing: 0.407 vs 0.405 ms - a predicated-off LDS / STS
// occupies the LSU pipe exactly like a predicated-on one (profiles/microbench/pred_lds.cu: 2.02 vs 2.04 cycles).  A 32-byte bin entry with w0 and
// the four byte offsets ready to use - 18 instead of 26 instructions per bin, but a second uniform LDS.128 - was 3 %
// SLOWER (0.417 ms): the kernel sits at 82 % of the LSU data pipe's wavefront rate, so a shared-memory wavefront is
// worth more than eight ALU instructions here.)
template <int PW, bool A, bool B>
__device__ __forceinline__ void tb_rmw2(const float (&hA)[PW], const float (&hB)[PW], const uint4* __restrict__ xt,
                                        char* rowA, float two_ic) {
#pragma unroll
  for (int pw = 0; pw < PW; ++pw) {
    const uint4 e = xt[pw];        // (two LDS.64 instead of one LDS.128: no difference, 0.407 vs 0.403 ms)
    const float w1 = __uint_as_float(e.x), w2 = __uint_as_float(e.y), w3 = __uint_as_float(e.z);
    const float w0 = two_ic - ((w1 + w2) + w3);
    float* p0 = reinterpret_cast<float*>(rowA + (e.w & 0xffu) * (kTbPix * 4));
    float* p1 = reinterpret_cast<float*>(rowA + ((e.w >> 8) & 0xffu) * (kTbPix * 4));
    float* p2 = reinterpret_cast<float*>(rowA + ((e.w >> 16) & 0xffu) * (kTbPix * 4));
    float* p3 = reinterpret_cast<float*>(rowA + (e.w >> 24) * (kTbPix * 4));
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
    if (A) { a0 = *p0; a1 = *p1; a2 = *p2; a3 = *p3; }
    if (B) { b0 = p0[kTbRowWords]; b1 = p1[kTbRowWords]; b2 = p2[kTbRowWords]; b3 = p3[kTbRowWords]; }
    if (A) {
      const float hv = hA[pw];
      *p0 = fmaf(w0, hv, a0); *p1 = fmaf(w1, hv, a1); *p2 = fmaf(w2, hv, a2); *p3 = fmaf(w3, hv, a3);
    }
    if (B) {
      const float hv = hB[pw];
      p0[kTbRowWords] = fmaf(w0, hv, b0); p1[kTbRowWords] = fmaf(w1, hv, b1);
      p2[kTbRowWords] = fmaf(w2, hv, b2); p3[kTbRowWords] = fmaf(w3, hv, b3);
    }
  }
}

template <int PW>
__global__ void __launch_bounds__(kTbThreads, MXD_TB_CTAS)
roi_align_tile_bwd_kernel(const __grid_constant__ FpnDesc d, const __grid_constant__ TCfg c, TWs w,
                          const float* __restrict__ gout) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* tile = reinterpret_cast<float*>(smem);
  unsigned char* stages = smem + c.tile_bytes;
  TbCtl* ctl = reinterpret_cast<TbCtl*>(smem + c.tile_bytes + (size_t)c.n_stages * c.stage_bytes);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < c.n_stages; ++s) {
      mbar_init(&ctl->full[s], 1);
      mbar_init(&ctl->empty[s], kTbWarps);
    }
    fence_mbar_init();
  }
  __syncthreads();
  if (warp >= kTbWarps) {
    pdl_wait();                     // the planner's lists and tables are complete
    tb_producer(d, c, w, gout, stages, ctl, lane, warp - kTbWarps);
    return;
  }
  // ================================= consumer warps ====================================
  // Warp w owns tile rows 2w and 2w+1 for the whole kernel (fixed pitch, independent of the level's tile shape): it
  // zeroes them once, accumulates every RoI of the item into them, and writes + re-zeroes them at the item's end.
  // No barrier ever: the only coupling between warps is the depth of the message ring.
  const int row0 = kTbR * warp;
  float* trow = tile + (size_t)(row0 < c.max_rows ? row0 : 0) * kTbRowWords;     // rows >= max_rows are never owned
  for (int r = 0; r < kTbR; ++r)
    if (row0 + r < c.max_rows)
      for (int i = lane; i < kTbRowWords; i += 32) trow[r * kTbRowWords + i] = 0.0f;
  __syncwarp();
  pdl_wait();                       // (consumers store to the gradient maps: everything before this launch is complete)
  pdl_launch_dependents();
  int s = 0;
  uint32_t par = 0;
  bool have = false, touchedA = false, touchedB = false;    // touched: some RoI of the item reached the row
  int lvl = 0, img = 0, c0 = 0, ty0 = 0, tx0 = 0, th = 0, tw = 0, H = 0, W = 0;
  const bool acc = c.accumulate != 0;
  // req=add: the first RoI that reaches a row loads the row's current gradient into the tile (coalesced, 8 loads in
  // flight per lane), so the write-out stays a plain store and untouched rows cost nothing.  (Adding on the way out -
  // a dependent global load per element inside the store loop - measured 1.03 ms against 0.41 ms for req=write.)
  auto load_row = [&](int r) {
    const int row = row0 + r;
    const int gy = ty0 + row;
    if (row >= th || gy >= H) return;
    float* tr = trow + r * kTbRowWords;
    const int twe = min(tw, W - tx0);
    const size_t plane = (size_t)H * W;
    const float* g0 = d.feat[lvl] + (((size_t)img * c.C + c0) * H + gy) * W + tx0;
    if (lane < twe) {
      const float* gp = g0 + lane;
      float* tp = tr + lane * kTbPix;
#pragma unroll 8
      for (int ch = 0; ch < 32; ++ch) { tp[ch] = __ldg(gp); gp += plane; }
    }
    const int x2 = 32 + (lane & 15), chb = (lane >> 4) * 16;
    if (x2 < twe) {
      const float* gp = g0 + (size_t)chb * plane + x2;
      float* tp = tr + x2 * kTbPix + chb;
#pragma unroll 8
      for (int ch = 0; ch < 16; ++ch) { tp[ch] = __ldg(gp); gp += plane; }
    }
    __syncwarp();
  };
  auto write_row = [&](int r, bool zeros) {
    const int row = row0 + r;
    const int gy = ty0 + row;
    if (row >= th || gy >= H) return;
    if (acc && zeros) return;               // req=add: nothing was added to this row
    float* tr = trow + r * kTbRowWords;
    const int twe = min(tw, W - tx0);
    const size_t plane = (size_t)H * W;
    float* g0 = d.feat[lvl] + (((size_t)img * c.C + c0) * H + gy) * W + tx0;
    if (lane < twe) {                       // columns 0..31: lane = column, one channel per step
      float* gp = g0 + lane;
      float* tp = tr + lane * kTbPix;
#pragma unroll 8
      for (int ch = 0; ch < 32; ++ch) {
        float v = 0.0f;
        if (!zeros) { v = tp[ch]; tp[ch] = 0.0f; }
        *gp = v;
        gp += plane;
      }
    }
    const int x2 = 32 + (lane & 15), chb = (lane >> 4) * 16;   // columns 32..47: two channels (ch, ch+16) per step
    if (x2 < twe) {
      float* gp = g0 + (size_t)chb * plane + x2;
      float* tp = tr + x2 * kTbPix + chb;
#pragma unroll 8
      for (int ch = 0; ch < 16; ++ch) {
        float v = 0.0f;
        if (!zeros) { v = tp[ch]; tp[ch] = 0.0f; }
        *gp = v;
        gp += plane;
      }
    }
    __syncwarp();
  };
#ifdef MXD_TB_PROF
  long long t_full = 0, t_row = 0, t_wr = 0, t_call = clock64();
#endif
  for (;;) {
#ifdef MXD_TB_PROF
    { const long long _t = clock64(); mbar_wait(&ctl->full[s], par); t_full += clock64() - _t; }
#else
    mbar_wait(&ctl->full[s], par);
#endif
    const unsigned char* st = stages + (size_t)s * c.stage_bytes;
    const int hd = reinterpret_cast<const int*>(st)[0];
    const int kind = hd & 0xff;
    if (kind == kMsgPair) {
      const int ra = (hd >> 8) & 0xff, rb = hd >> 16;
      const bool cA = row0 >= ra && row0 <= rb, cB = kTbR > 1 && row0 + 1 >= ra && row0 + 1 <= rb;
      if (cA || cB) {
#ifdef MXD_TB_PROF
        const long long _t = clock64();
#endif
        const float* gl = reinterpret_cast<const float*>(st + c.off_g) + lane * c.bins;
        const uint2* rt = reinterpret_cast<const uint2*>(st + c.off_rt) + row0 * 4;
        const uint4* xt = reinterpret_cast<const uint4*>(st + c.off_xt);
        char* rowp = reinterpret_cast<char*>(trow + lane);
        if constexpr (kTbR == 1) {
          float hA[PW];
          if (tb_h<PW>(gl, rt, hA)) {
            if (acc && !touchedA) load_row(0);
            touchedA = true;
            tb_rmw2<PW, true, false>(hA, hA, xt, rowp, 2.0f * c.inv_count);
          }
        } else {
          float hA[PW], hB[PW];
          const bool fA = cA && tb_h<PW>(gl, rt, hA), fB = cB && tb_h<PW>(gl, rt + 4, hB);
          if (acc) {
            if (fA && !touchedA) load_row(0);
            if (fB && !touchedB) load_row(1);
          }
          touchedA = touchedA || fA; touchedB = touchedB || fB;
          if (fA && fB) tb_rmw2<PW, true, true>(hA, hB, xt, rowp, 2.0f * c.inv_count);
          else if (fA) tb_rmw2<PW, true, false>(hA, hB, xt, rowp, 2.0f * c.inv_count);
          else if (fB) tb_rmw2<PW, false, true>(hA, hB, xt, rowp, 2.0f * c.inv_count);
        }
#ifdef MXD_TB_PROF
        t_row += clock64() - _t;
#endif
      }
    } else {
#ifdef MXD_TB_PROF
      const long long _t = clock64();
#endif
      if (have) { write_row(0, !touchedA); if (kTbR > 1) write_row(1, !touchedB); }     // an untouched row is still all zero: store zeros only
#ifdef MXD_TB_PROF
      t_wr += clock64() - _t;
      if (kind == kMsgStop && lane == 0) {
        atomicAdd(reinterpret_cast<unsigned long long*>(w.hdr + 20), (unsigned long long)t_full);
        atomicAdd(reinterpret_cast<unsigned long long*>(w.hdr + 22), (unsigned long long)t_row);
        atomicAdd(reinterpret_cast<unsigned long long*>(w.hdr + 24), (unsigned long long)t_wr);
        atomicAdd(reinterpret_cast<unsigned long long*>(w.hdr + 26), (unsigned long long)(clock64() - t_call));
      }
#endif
      have = false; touchedA = touchedB = false;
      if (kind == kMsgStop) break;
      const int4 h1 = reinterpret_cast<const int4*>(st)[1];
      lvl = h1.x; img = h1.y; c0 = h1.z; ty0 = h1.w & 0xffff; tx0 = h1.w >> 16;
      th = c.lv[lvl].th; tw = c.lv[lvl].tw; H = c.lv[lvl].H; W = c.lv[lvl].W;
      if (kind == kMsgBegin) have = true;
      else { write_row(0, true); if (kTbR > 1) write_row(1, true); }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&ctl->empty[s]);
    if (++s == c.n_stages) { s = 0; par ^= 1u; }
  }
}

// RoIs the planner rejected: generic per-tap RED path, after the tiles are in HBM.
__global__ void __launch_bounds__(256) tile_bwd_fallback_kernel(FpnDesc d, TCfg c, TWs w,
                                                                 const float* __restrict__ rois,
                                                                 const int* __restrict__ levels, float* gout) {
  pdl_wait();                       // every tile is in HBM
  const int n_fb = w.hdr[1];
  const int chunks = (c.C + 31) / 32;
  for (int i = blockIdx.x; i < n_fb * chunks; i += gridDim.x) {
    const int n = w.fb_list[i / chunks];
    const int c0 = (i % chunks) * 32;
    const RoiGeom g = roi_geom(d, rois, levels, n, c.PH, c.PW, c.sr, c.finest);
    gather_roi_chunk<true>(d, g, n, c0, min(32, c.C - c0), gout, c.PH, c.PW, threadIdx.x, 256);
  }
}

#ifdef MXD_TB_PROF
extern "C" int mxd_tb_prof(const void* ws, unsigned long long* out6) {
  return (int)cudaMemcpy(out6, (const char*)ws + 64, 48, cudaMemcpyDeviceToHost);
}
#endif

size_t tile_bwd_workspace_bytes(int R, int N, int L, const int* Hs, const int* Ws, int C, int PH, int PW, int sr) {
  TCfg c;
  if (!make_tcfg(N, C, L, Hs, Ws, PH, PW, sr, 56.0f, 0, &c)) return 0;
  return carve_tile(nullptr, R, c.NT, c.tx, c.max_hf).bytes;
}

int tile_backward(const FpnDesc& d, const float* rois, const int* levels, const float* gout, int R, int PH, int PW,
                  int sr, float finest, int accumulate, void* ws, size_t ws_bytes, cudaStream_t st, int* handled) {
  *handled = 0;
  TCfg c;
  if (R == 0 || d.C == 0) return MXD_OK;
  if (!make_tcfg(d.N, d.C, d.num_levels, d.H, d.W, PH, PW, sr, finest, accumulate, &c)) return MXD_OK;
  if ((reinterpret_cast<uintptr_t>(gout) & 15) != 0) return MXD_OK;     // TMA source alignment
  TWs w = carve_tile(ws, R, c.NT, c.tx, c.max_hf);
  MXD_REQUIRE(ws_bytes >= w.bytes, MXD_EWORKSPACE, "roi_align workspace %zu < %zu bytes", ws_bytes, w.bytes);
  MXD_REQUIRE(((uintptr_t)ws & 255) == 0, MXD_EINVAL, "workspace must be 256-byte aligned");
  int sms = 0, dev = 0;
  MXD_CUDA_OK(cudaGetDevice(&dev));
  MXD_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t zbytes = (size_t)((char*)w.start - (char*)w.hdr);    // hdr, cnt, cursor
  MXD_CUDA_OK(cudaMemsetAsync(w.hdr, 0, zbytes, st));
  if (PH == 7) MXD_CUDA_OK(launch_pdl(tplan_rois_kernel<7>, dim3((R * 32 + 255) / 256), dim3(256), 0, st, d, c, w, rois, levels, R));
  else MXD_CUDA_OK(launch_pdl_if(false, tplan_rois_kernel<14>, dim3((R * 32 + 255) / 256), dim3(256), 0, st, d, c, w, rois, levels, R));
  MXD_POST_LAUNCH("roi_align_tplan_rois");
  // (the 14x14 chain measured 37 us SLOWER with programmatic launches - 0.419 vs 0.382 ms on BASELINE config 4a - and keeps
  // plain stream order; the 7x7 chain gains 9 us)
  const bool pdl = PW == 7;
  MXD_CUDA_OK(launch_pdl_if(pdl, tplan_group_kernel, dim3(1), dim3(1024), 0, st, c, w, R));
  MXD_POST_LAUNCH("roi_align_tplan_group");
  MXD_CUDA_OK(launch_pdl_if(pdl, tplan_sort_kernel, dim3((c.NT + 7) / 8), dim3(256), 0, st, c, w));
  MXD_POST_LAUNCH("roi_align_tplan_sort");
  static unsigned long long seen = 0;
  DeviceOnce once_seen(&seen);
  if (once_seen.first()) {
    MXD_CUDA_OK(cudaFuncSetAttribute(roi_align_tile_bwd_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTbSmem));
    MXD_CUDA_OK(cudaFuncSetAttribute(roi_align_tile_bwd_kernel<14>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTbSmem));
  }
  if (PW == 7) MXD_CUDA_OK(launch_pdl(roi_align_tile_bwd_kernel<7>, dim3(sms * MXD_TB_CTAS), dim3(kTbThreads), (size_t)c.smem_bytes, st, d, c, w, gout));
  else MXD_CUDA_OK(launch_pdl_if(false, roi_align_tile_bwd_kernel<14>, dim3(sms * MXD_TB_CTAS), dim3(kTbThreads), (size_t)c.smem_bytes, st, d, c, w, gout));
  MXD_POST_LAUNCH("roi_align_tile_bwd");
  MXD_CUDA_OK(launch_pdl_if(pdl, tile_bwd_fallback_kernel, dim3(2 * sms), dim3(256), 0, st, d, c, w, rois, levels, const_cast<float*>(gout)));
  MXD_POST_LAUNCH("roi_align_tile_bwd_fallback");
  *handled = 1;
  return MXD_OK;
}

}  // namespace mxd
