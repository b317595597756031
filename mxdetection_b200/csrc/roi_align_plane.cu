// Plane-resident RoIAlign forward / backward for B200 (Spec A, G).
//
// Contract as roi_align.cu (mx.nd.contrib.ROIAlign, mxdetection/ops,
// /root/reference/README.md:24; SingleLevelRoI.forward,
// /root/reference/README.md:32).  Design:
//
//  * NCHW makes one channel plane a contiguous byte range, so a band of rows of a
//    plane is ONE cp.async.bulk (TMA, UBLKCP) into shared memory, completion
//    signalled on an mbarrier; no register staging, no LSU instructions.
//  * A planner kernel (one warp per RoI) evaluates Spec A's per-axis sample
//    tables once per RoI (they are shared by all C channels), assigns the RoI to
//    the row band holding its first tap row and packs the tables to 8 B / entry.
//  * A persistent kernel (one or two CTAs per SM, work-stealing over an atomic
//    counter) walks items (image, level, band, channel group): it lands the band
//    in shared memory, then every thread produces outputs (RoI, bin) of that band
//    with 16 shared-memory taps per output; each plane byte crosses L2->SM once
//    per band instead of once per RoI, and never leaves L2 twice.
//  * RoIs that do not fit a band window, sample outside the image or have a bad
//    batch index are few; the same kernel processes them afterwards with the
//    generic global-memory gather (gather_roi_chunk).
#include "roi_align.cuh"
#include "ptx.cuh"

namespace mxd {

constexpr int kSmemLimit = 227 * 1024;       // per-CTA opt-in maximum on sm_100
constexpr int kBigRing = 14 * 1024;          // table area in one-CTA-per-SM mode: leaves 54400 floats = a 200x272 plane
constexpr int kSmallRing = 8 * 1024;
constexpr int kMiscBytes = 512;              // mbarrier + scalars + staged RoI ids
// stream forward kernel (7x7 / sample_ratio 2): two band buffers + two table buffers per CTA
constexpr int kS4TabBytes = 16 * 1024;       // one table buffer
// grouped table of one RoI: ty + tx tap entries + {RoI id, 0}, padded to an even count (16-byte multiple)
static inline int s4_ent(int entries) { return (entries + 2) & ~1; }
constexpr int kS4KC = 4;                     // items (channel groups) per work unit
constexpr int kS4CtlBytes = 512;
constexpr int kS4Bufs = 2;                   // band buffers of the stream kernel (3 x 66 KB measured 8 % slower:
                                             // shorter level-0 windows, level 1 no longer a whole plane)

struct PlanLevel {
  int H, W;
  int band_rows;   // anchor band height (0: level not plane-capable -> gather fallback)
  int max_rows;    // rows of one window that fit the smem budget
  int nbands;
  int cg, ncg;     // channels per item, number of channel groups
  int band_base;   // first band of this level inside one image's band table
  int item_base;   // first item id of this level
  int n_items;     // N * nbands * ncg
  int kc, nchunk;  // stream kernel: items (channel groups) per work unit, chunks per band
  int unit_base;   // stream kernel: first unit id of this level
};

struct PlanCfg {
  PlanLevel lv[MXD_MAX_LEVELS];
  int L, N, C, PH, PW, sr, ty, tx;
  int bands_per_img, NB, n_plane_items;
  int budget_floats, chunk_rois, tab_bytes, slot_bytes;
  int threads, smem_bytes, ctas_per_sm, group;
  int stream, n_units, s4_ent, s4_max_rois;
  float finest;
};

struct PlanWs {
  int* hdr;      // [0] work counter, [1] fallback count
  int* cnt;      // [NB] RoIs per band
  int* rmax;     // [NB] last tap row over the band's RoIs
  int* start;    // [NB] exclusive prefix of cnt
  int* cursor;   // [NB]
  int* meta;     // [R] band index or -1
  int* list;     // [R] RoI ids grouped by band
  int* fb_list;  // [R] RoIs for the gather fallback
  uint2* tab;    // [R][ty+tx] packed {offset | hi_bit<<31, lo-weight bits}
  uint2* tabg;   // [R][s4_ent] the same tables in band-grouped order + {RoI id, 0} (stream kernel)
  size_t bytes;
};

static PlanWs carve_plan(void* base, int R, int NB, int entries) {
  PlanWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return (char*)base + o; };
  w.hdr = (int*)take(sizeof(int) * 64);
  w.cnt = (int*)take(sizeof(int) * (size_t)NB);
  w.rmax = (int*)take(sizeof(int) * (size_t)NB);
  w.start = (int*)take(sizeof(int) * (size_t)NB);
  w.cursor = (int*)take(sizeof(int) * (size_t)NB);
  w.meta = (int*)take(sizeof(int) * (size_t)(R > 0 ? R : 1));
  w.list = (int*)take(sizeof(int) * (size_t)(R > 0 ? R : 1));
  w.fb_list = (int*)take(sizeof(int) * (size_t)(R > 0 ? R : 1));
  w.tab = (uint2*)take(sizeof(uint2) * (size_t)(R > 0 ? R : 1) * entries);
  w.tabg = (uint2*)take(sizeof(uint2) * (size_t)(R > 0 ? R : 1) * s4_ent(entries));
  w.bytes = off;
  return w;
}

// Host: choose bands / channel groups per level for the shared-memory budget.
static bool make_cfg(int N, int C, int L, const int* Hs, const int* Ws, int PH, int PW, int sr, float finest,
                     PlanCfg* c, bool stream = false) {
  if (sr <= 0 || PH * sr > 64 || PW * sr > 32) return false;   // adaptive / very fine sampling: gather kernels
  if (stream && (sr != 2 || !((PH == 7 && PW == 7) || (PH == 14 && PW == 14)))) return false;
  memset(c, 0, sizeof(*c));
  c->stream = stream ? 1 : 0;
  c->s4_ent = s4_ent(PH * sr + PW * sr);
  c->s4_max_rois = kS4TabBytes / (c->s4_ent * 8);
  c->L = L; c->N = N; c->C = C; c->PH = PH; c->PW = PW; c->sr = sr; c->ty = PH * sr; c->tx = PW * sr;
  c->finest = finest;
  const int entry_bytes = (c->ty + c->tx) * (int)sizeof(uint2);
  // "big" mode: one CTA per SM with (almost) all shared memory, when some plane fits only that way
  const int small_est = (kSmemLimit / 2 - 1024 - kSmallRing - kMiscBytes) / 4;
  const int big_est = (kSmemLimit - kBigRing - kMiscBytes) / 4;
  bool big = false;
  for (int l = 0; l < L; ++l) {
    const size_t plane = (size_t)Hs[l] * Ws[l];
    if (plane > (size_t)small_est && plane <= (size_t)big_est) big = true;
  }
  // planes that fit neither way are banded; banding prefers two CTAs per SM
  for (int l = 0; l < L; ++l)
    if ((size_t)Hs[l] * Ws[l] > (size_t)big_est) big = false;
  // one (RoI, channel) job per lane group (16 lanes when PW*sr <= 16, else a warp)
  c->group = (c->tx <= 16) ? 16 : 32;
  c->ctas_per_sm = big ? 1 : 2;
  const int consumers = big ? 992 : 480;           // + one producer warp
  c->threads = consumers + 32;
  // every lane group owns one private slot for the packed tap table of its current RoI
  const int smem_cta = big ? kSmemLimit : (kSmemLimit / 2 - 1024);
  c->chunk_rois = consumers / c->group;
  c->tab_bytes = (int)align_up((size_t)c->chunk_rois * entry_bytes, 128);
  c->slot_bytes = entry_bytes;
  if (2 * c->group < c->ty + c->tx) return false;   // a lane moves at most two table entries
  c->budget_floats = ((smem_cta - c->tab_bytes - kMiscBytes) / 4) & ~3;
  if (stream) {
    if (big) return false;                 // a plane that only fits whole: the one-CTA plane kernel keeps it resident
    c->ctas_per_sm = 1; c->threads = (4 * PW <= 32) ? 1024 : 768; c->tab_bytes = 2 * kS4TabBytes;
    c->budget_floats = (((kSmemLimit - 2 * kS4TabBytes - kS4CtlBytes) / kS4Bufs) / 4) & ~3;
  }
  if (c->budget_floats < 4096) return false;
  c->smem_bytes = stream ? kS4Bufs * c->budget_floats * 4 + 2 * kS4TabBytes + kS4CtlBytes
                         : c->budget_floats * 4 + c->tab_bytes + kMiscBytes;
  int band_base = 0, item_base = 0;
  for (int l = 0; l < L; ++l) {
    PlanLevel& v = c->lv[l];
    v.H = Hs[l]; v.W = Ws[l];
    const int fit_rows = c->budget_floats / v.W;
    if (fit_rows >= v.H) {                       // whole plane(s) resident
      v.max_rows = v.H; v.band_rows = v.H; v.nbands = 1;
      v.cg = c->budget_floats / (v.H * v.W);
      if (v.cg > C) v.cg = C;
      if (v.cg > 32) v.cg = 32;
    } else if (fit_rows >= 16) {                 // row bands: anchor bands of ~60 % of the window
      v.max_rows = fit_rows;
      v.band_rows = (fit_rows * 3) / 5;
      v.nbands = (v.H + v.band_rows - 1) / v.band_rows;
      v.cg = 1;
    } else {                                      // rows too wide for a useful window
      v.max_rows = 0; v.band_rows = 0; v.nbands = 0; v.cg = 1;
    }
    v.ncg = (C + v.cg - 1) / v.cg;
    v.band_base = band_base; band_base += v.nbands;
    v.item_base = item_base; v.n_items = N * v.nbands * v.ncg; item_base += v.n_items;
    if (stream) {     // every level must be plane/band capable and loadable with 16-byte-granular bulk copies
      if (v.band_rows == 0) return false;
      if (v.nbands == 1) {          // whole planes: one contiguous copy per channel group
        const size_t pb = (size_t)v.H * v.W * 4;
        int al = 1;
        while ((pb * al) & 15) al <<= 1;
        if (v.cg < al || (C % al) != 0) return false;
        if (v.cg > 8) v.cg = 8;     // keep work units of all levels comparable (dynamic scheduling tail)
        v.cg = v.cg / al * al;
        if (v.cg < al) v.cg = al;
        v.ncg = (C + v.cg - 1) / v.cg;
        v.n_items = N * v.nbands * v.ncg;
      } else if ((v.W & 3) != 0) {
        return false;
      }
    }
    v.kc = kS4KC / v.cg > 1 ? kS4KC / v.cg : 1;
    v.nchunk = (v.ncg + v.kc - 1) / v.kc;
  }
  for (int l = L - 1; l >= 0; --l) {      // unit ids: coarsest level first - its units carry the most RoIs
    PlanLevel& v = c->lv[l];
    v.unit_base = c->n_units; c->n_units += N * v.nbands * v.nchunk;
  }
  c->bands_per_img = band_base;
  c->NB = N * band_base;
  c->n_plane_items = item_base;
  return c->NB > 0;
}

// ------------------------------------------------------------------- planner ------
__global__ void __launch_bounds__(256) plan_rois_kernel(FpnDesc d, PlanCfg c, PlanWs w,
                                                         const float* __restrict__ rois,
                                                         const int* __restrict__ levels, int R, int bwd) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= R) return;
  const RoiGeom g = roi_geom(d, rois, levels, n, c.PH, c.PW, c.sr, c.finest);
  const PlanLevel& v = c.lv[g.ok ? g.lvl : 0];
  bool ok = g.ok && (bwd || v.band_rows > 0) && g.H >= 2 && g.W >= 2;
  AxisTap ya[2], xa[2];
  int rfirst = 0x7fffffff, rlast = -1;
  bool valid = true;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int t = lane + 32 * k;
    if (t < c.ty) {
      ya[k] = axis_tap(g.rsh, g.bh, c.sr, t / c.sr, t % c.sr, g.H, 1);
      valid = valid && ya[k].valid;
      rfirst = min(rfirst, ya[k].hi == ya[k].lo ? ya[k].lo - 1 : ya[k].lo); rlast = max(rlast, ya[k].hi);
    }
    if (t < c.tx) {
      xa[k] = axis_tap(g.rsw, g.bw, c.sr, t / c.sr, t % c.sr, g.W, 1);
      valid = valid && xa[k].valid;
    }
  }
  rfirst = __reduce_min_sync(0xffffffffu, rfirst);
  rlast = __reduce_max_sync(0xffffffffu, rlast);
  ok = ok && __all_sync(0xffffffffu, valid);
  int band = 0, r0 = 0;
  if (ok && !bwd) {
    band = rfirst / v.band_rows;
    r0 = band * v.band_rows;
    if (rlast - r0 + 1 > v.max_rows) ok = false;   // the whole footprint must sit inside one window
  }
  if (!ok) {
    if (lane == 0) {
      w.meta[n] = -1;
      if (!bwd) w.fb_list[atomicAdd(&w.hdr[1], 1)] = n;   // backward: the rows kernel checks meta itself
    }
    return;
  }
  const int bidx = g.b * c.bands_per_img + v.band_base + band;
  uint2* tab = w.tab + (size_t)n * (c.ty + c.tx);
  // Entries are stored in "unclamped" form: low tap t, high tap t+1, weight l of the high tap.  A sample
  // clamped at the border (lo == hi == size-1, l == 0) becomes (size-2, l = 1): h*D[size-2] + l*D[size-1]
  // = D[size-1], the same value, so the kernel can address the 2x2 patch with fixed +1 / +pitch offsets.
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int t = lane + 32 * k;
    if (t < c.ty) {
      const bool cl = ya[k].hi == ya[k].lo;
      const int lo = cl ? ya[k].lo - 1 : ya[k].lo;
      // forward: byte offset of the low row inside the window; backward: absolute low row
      const unsigned off = bwd ? (unsigned)lo : (unsigned)((lo - r0) * v.W * 4);
      tab[t] = make_uint2(off, __float_as_uint(cl ? 1.0f : ya[k].l));
    }
    if (t < c.tx) {
      const bool cl = xa[k].hi == xa[k].lo;
      tab[c.ty + t] = make_uint2((unsigned)((cl ? xa[k].lo - 1 : xa[k].lo) * 4), __float_as_uint(cl ? 1.0f : xa[k].l));
    }
  }
  if (lane == 0) {
    if (!bwd) {
      w.meta[n] = bidx;
      atomicAdd(&w.cnt[bidx], 1);
      atomicMax(&w.rmax[bidx], rlast);
    } else {
      w.meta[n] = 0;     // backward: tables only (rows kernel), no band lists
    }
  }
}

// Exclusive scan of the band counts + grouping of the RoI ids (single CTA).
__global__ void __launch_bounds__(1024) plan_group_kernel(PlanCfg c, PlanWs w, int R, int bwd) {
  __shared__ int s_part[1024];
  const int tid = threadIdx.x;
  const int per = (c.NB + 1023) / 1024;
  int sum = 0;
  for (int i = tid * per; i < min(c.NB, (tid + 1) * per); ++i) sum += w.cnt[i];
  s_part[tid] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    int v = tid >= o ? s_part[tid - o] : 0;
    __syncthreads();
    s_part[tid] += v;
    __syncthreads();
  }
  int run = s_part[tid] - sum;
  for (int i = tid * per; i < min(c.NB, (tid + 1) * per); ++i) {
    w.start[i] = run;
    run += w.cnt[i];
  }
  __syncthreads();
  __threadfence_block();
  for (int n = tid; n < R; n += 1024) {
    const int m = w.meta[n];
    if (m < 0) continue;
    w.list[w.start[m] + atomicAdd(&w.cursor[m], 1)] = n;
  }
}

// Stream kernel: the tables in band-grouped order, so one bulk copy brings a band's RoIs to shared memory.
__global__ void __launch_bounds__(256) plan_pack_kernel(PlanCfg c, PlanWs w, int R) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int pos = i / c.s4_ent, e = i - pos * c.s4_ent;
  const int ent = c.ty + c.tx;
  if (pos >= R - w.hdr[1]) return;                // planned RoIs = all but the fallback ones
  const int n = w.list[pos];
  uint2 v = make_uint2(0u, 0u);
  if (e < ent) v = w.tab[(size_t)n * ent + e];
  else if (e == ent) v = make_uint2((unsigned)n, 0u);
  w.tabg[(size_t)pos * c.s4_ent + e] = v;
}

__device__ __forceinline__ float ldf(const char* p) { return *reinterpret_cast<const float*>(p); }

struct ItemDesc { int lvl, img, band, cgi; };

__device__ __forceinline__ ItemDesc decode_item(const PlanCfg& c, int item) {
  ItemDesc it;
  int l = 0;
  while (l + 1 < c.L && item >= c.lv[l + 1].item_base) ++l;
  const PlanLevel& v = c.lv[l];
  int r = item - v.item_base;
  it.lvl = l;
  it.cgi = r % v.ncg; r /= v.ncg;
  it.band = r % v.nbands;
  it.img = r / v.nbands;
  return it;
}

// --------------------------------------------------------------- forward ---------
// Shared-memory map of the persistent kernels:
//   [band buffer: budget_floats*4][per-group tap tables: ngroups*ent*8][SmemCtl]
struct ItemSlot {
  int kind;            // 0 plane item, 1 gather fallback unit, 2 stop
  int lvl, img, c0, ncur, cnt, nrows, bulk;
  int chan_bytes, pitch_bytes;
  int lst;             // first entry of the band's RoI list
  int r0;              // first row of the window
  int fb_roi, fb_c0;
};
struct SmemCtl {
  u64 desc_full, band_full, band_empty;
  ItemSlot item;
};


// Producer warp of the persistent kernels: pulls items off the global counter, waits until the
// consumers released the band buffer, publishes the item descriptor and issues the TMA band loads.
__device__ __forceinline__ void plane_producer(const FpnDesc& d, const PlanCfg& c, const PlanWs& w, SmemCtl* ctl,
                                               float* buf, int lane, bool load_band) {
  const int n_fb = w.hdr[1];
  const int fb_chunks = (c.C + 31) / 32;
  // the backward runs its fallback RoIs in a second kernel (after the bands are written)
  const int total = c.n_plane_items + (load_band ? n_fb * fb_chunks : 0);
  uint32_t band_phase = 0;
  for (;;) {
    int item = 0;
    if (lane == 0) item = atomicAdd(&w.hdr[0], 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    ItemSlot it;
    it.kind = 2;
    it.lvl = it.img = it.c0 = it.ncur = it.cnt = it.nrows = it.bulk = it.chan_bytes = it.pitch_bytes = 0;
    it.lst = it.r0 = it.fb_roi = it.fb_c0 = 0;
    const float* src0 = nullptr;
    size_t plane_sz = 0;
    if (item < c.n_plane_items) {
      const ItemDesc de = decode_item(c, item);
      const PlanLevel& v = c.lv[de.lvl];
      const int bidx = de.img * c.bands_per_img + v.band_base + de.band;
      const int cnt = w.cnt[bidx];
      if (cnt == 0 && load_band) continue;   // forward: nothing to pool from this band (warp-uniform)
      it.r0 = de.band * v.band_rows;
      it.kind = 0; it.lvl = de.lvl; it.img = de.img; it.cnt = cnt;
      it.nrows = load_band ? min(w.rmax[bidx], v.H - 1) - it.r0 + 1 : min(v.band_rows, v.H - it.r0);
      it.c0 = de.cgi * v.cg;
      it.ncur = min(v.cg, c.C - it.c0);
      it.pitch_bytes = v.W * 4;
      it.chan_bytes = it.nrows * v.W * 4;
      it.bulk = load_band && ((v.W & 3) == 0) && ((reinterpret_cast<uintptr_t>(d.feat[de.lvl]) & 15) == 0);
      it.lst = w.start[bidx];
      src0 = d.feat[de.lvl] + (((size_t)de.img * c.C + it.c0) * v.H + it.r0) * v.W;
      plane_sz = (size_t)v.H * v.W;
    } else if (item < total) {
      const int f = item - c.n_plane_items;
      it.kind = 1; it.fb_roi = w.fb_list[f / fb_chunks]; it.fb_c0 = (f % fb_chunks) * 32;
    }
    // the band buffer (and the descriptor) is free once every consumer warp released it
    mbar_wait(&ctl->band_empty, band_phase ^ 1);
    if (lane == 0) {
      ctl->item = it;
      mbar_arrive(&ctl->desc_full);
      if (it.kind == 0 && it.bulk) {
        fence_proxy_async();
        mbar_arrive_expect_tx(&ctl->band_full, (uint32_t)(it.ncur * it.chan_bytes));
        for (int j = 0; j < it.ncur; ++j)
          bulk_g2s(reinterpret_cast<char*>(buf) + (size_t)j * it.chan_bytes, src0 + (size_t)j * plane_sz,
                   (uint32_t)it.chan_bytes, &ctl->band_full);
      } else {
        mbar_arrive(&ctl->band_full);
      }
    }
    band_phase ^= 1;
    if (it.kind == 2) break;
  }
}

template <int SR>
__global__ void __launch_bounds__(1024, 1)
roi_align_plane_fwd_kernel(const __grid_constant__ FpnDesc d, const __grid_constant__ PlanCfg c, PlanWs w,
                           const float* __restrict__ rois, const int* __restrict__ levels,
                           float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* buf = reinterpret_cast<float*>(smem);
  uint2* tab_all = reinterpret_cast<uint2*>(smem + (size_t)c.budget_floats * 4);
  SmemCtl* ctl = reinterpret_cast<SmemCtl*>(smem + (size_t)c.budget_floats * 4 + c.tab_bytes);
  const int tid = threadIdx.x;
  const int Tc = blockDim.x - 32;            // consumer threads; the last warp is the producer
  const int bins = c.PH * c.PW;
  const int ent = c.ty + c.tx;
  const int sr = SR > 0 ? SR : c.sr;
  const int n_cwarps = Tc >> 5;
  if (tid == 0) {
    mbar_init(&ctl->desc_full, 1);
    mbar_init(&ctl->band_full, 1);
    mbar_init(&ctl->band_empty, n_cwarps);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid >= Tc) {
    plane_producer(d, c, w, ctl, buf, tid - Tc, true);
    return;
  }

  // ================================= consumer warps ====================================
  // One (RoI, channel) job per lane group (c.group = 16 or 32 lanes).  Lane gl owns x-sample gl:
  // consecutive lanes read increasing x of one feature row (bank-conflict free); the 2x2 patch of a
  // sample sits at [R], [R+4], [R+pitch], [R+pitch+4] with R = plane + x_off + y_off; the sr x-samples of
  // a bin are summed with shuffles and the lane of the first one stores the bin.  Each group keeps its
  // RoI's packed tap table in a private shared-memory slot, prefetched through registers one RoI ahead
  // (ids two ahead) so no global latency sits between jobs.
  const int grp = tid / c.group, gl = tid % c.group, ngroups = Tc / c.group;
  uint2* te = tab_all + (size_t)grp * ent;
  const float inv_count = 1.0f / (float)(sr * sr);
  const bool warp_lead = (tid & 31) == 0;
  const int e0 = gl, e1 = gl + c.group;      // table entries this lane moves (ent <= 2*group)
  uint32_t phase = 0;
  for (;;) {
    mbar_wait(&ctl->desc_full, phase);
    const ItemSlot it = ctl->item;
    if (it.kind == 2) break;
    if (it.kind == 1) {
      mbar_wait(&ctl->band_full, phase);
      phase ^= 1;
      const RoiGeom g = roi_geom(d, rois, levels, it.fb_roi, c.PH, c.PW, c.sr, c.finest);
      gather_roi_chunk<false>(d, g, it.fb_roi, it.fb_c0, min(32, c.C - it.fb_c0), out, c.PH, c.PW, tid, Tc);
      __syncwarp();
      if (warp_lead) mbar_arrive(&ctl->band_empty);
      continue;
    }
    // prefetch ids / first table while the band is in flight
    const int* list = w.list + it.lst;
    int id0 = grp < it.cnt ? list[grp] : -1;
    int id1 = grp + ngroups < it.cnt ? list[grp + ngroups] : -1;
    uint2 t0 = make_uint2(0, 0), t1 = make_uint2(0, 0);
    if (id0 >= 0) {
      if (e0 < ent) t0 = w.tab[(size_t)id0 * ent + e0];
      if (e1 < ent) t1 = w.tab[(size_t)id0 * ent + e1];
    }
    mbar_wait(&ctl->band_full, phase);
    phase ^= 1;
    if (!it.bulk) {   // rows not 16-byte aligned: cooperative load by the consumers
      const PlanLevel& v = c.lv[it.lvl];
      const int chan_floats = it.chan_bytes >> 2;
      const float* src0 = d.feat[it.lvl] + (((size_t)it.img * c.C + it.c0) * v.H + it.r0) * v.W;
      for (int i = tid; i < it.ncur * chan_floats; i += Tc) {
        const int j = i / chan_floats;
        buf[i] = __ldg(src0 + (size_t)j * v.H * v.W + (i - j * chan_floats));
      }
      consumer_sync(Tc);
    }
    const int pitch_bytes = it.pitch_bytes;
    const int rounds = (it.cnt + ngroups - 1) / ngroups;
    for (int k = 0; k < rounds; ++k) {
      const int n = id0;
      const bool live = n >= 0;
      __syncwarp();                       // previous job's reads of the slot are done
      if (e0 < ent) te[e0] = t0;
      if (e1 < ent) te[e1] = t1;
      __syncwarp();
      {   // prefetch: id two jobs ahead, table one job ahead
        const int q2 = grp + (k + 2) * ngroups;
        const int id2 = q2 < it.cnt ? list[q2] : -1;
        if (id1 >= 0) {
          if (e0 < ent) t0 = w.tab[(size_t)id1 * ent + e0];
          if (e1 < ent) t1 = w.tab[(size_t)id1 * ent + e1];
        }
        id0 = id1; id1 = id2;
      }
      const bool lane_on = live && gl < c.tx;
      const uint2 ex = te[c.ty + (gl < c.tx ? gl : 0)];
      const float lx = __uint_as_float(ex.y), hx = 1.0f - lx;
      const bool writer = lane_on && (gl % sr == 0);
      for (int j = 0; j < it.ncur; ++j) {
        const char* px = reinterpret_cast<const char*>(buf) + (size_t)j * it.chan_bytes + (live ? ex.x : 0);
        float* o = out + ((size_t)(live ? n : 0) * c.C + it.c0 + j) * bins + gl / sr;
        for (int ph = 0; ph < c.PH; ++ph) {
          float acc = 0.0f;
#pragma unroll
          for (int iy = 0; iy < (SR > 0 ? SR : 1); ++iy) {
            const uint2 ey = te[ph * sr + iy];
            const char* r = px + (live ? ey.x : 0);
            const float ly = __uint_as_float(ey.y), hy = 1.0f - ly;
            const float a = hx * ldf(r) + lx * ldf(r + 4);
            const float b = hx * ldf(r + pitch_bytes) + lx * ldf(r + pitch_bytes + 4);
            acc += hy * a + ly * b;
          }
          if (SR == 0) {   // runtime sample ratio: remaining rows
            for (int iy = 1; iy < sr; ++iy) {
              const uint2 ey = te[ph * sr + iy];
              const char* r = px + (live ? ey.x : 0);
              const float ly = __uint_as_float(ey.y), hy = 1.0f - ly;
              const float a = hx * ldf(r) + lx * ldf(r + 4);
              const float b = hx * ldf(r + pitch_bytes) + lx * ldf(r + pitch_bytes + 4);
              acc += hy * a + ly * b;
            }
          }
          float sum;
          if (SR == 2) {
            sum = acc + __shfl_xor_sync(0xffffffffu, acc, 1);
          } else {
            const int base = (threadIdx.x & 31) - gl % sr;
            sum = 0.0f;
            for (int kk = 0; kk < sr; ++kk) sum += __shfl_sync(0xffffffffu, acc, base + kk);
          }
          if (writer) o[ph * c.PW] = sum * inv_count;
        }
      }
    }
    __syncwarp();
    if (warp_lead) mbar_arrive(&ctl->band_empty);
  }
}

// ------------------------------------------------- forward, 7x7 / sample_ratio 2 ---
// Specialisation for the detector's bbox RoIAlign (2 * PW * sr = 28 x taps <= 32 lanes).  Measured on the
// generic kernel above: two 16-lane jobs per warp read different rows in one LDS and collide on banks
// (1.7 wavefronts per LDS), and the LSU data pipe - not issue - bounds the kernel.  Here ONE (RoI, channel)
// job owns the warp and lane l is x TAP l (sample l>>1, low/high tap l&1): all 28 taps of a sample row are
// one conflict-free LDS (<= 32 consecutive columns of one row), lane weights are per-lane registers, and the
// 4 lanes of a bin are folded with a 6-shuffle transposing butterfly that leaves every lane with at most
// two finished bins to store.
template <int PH, int PW>
__global__ void __launch_bounds__(1024, 1)
roi_align_plane_fwd_tap_kernel(const __grid_constant__ FpnDesc d, const __grid_constant__ PlanCfg c, PlanWs w,
                               const float* __restrict__ rois, const int* __restrict__ levels,
                               float* __restrict__ out) {
  static_assert(4 * PW <= 32 && PH <= 8, "lane = x tap, 8 accumulators");
  constexpr int TY = 2 * PH, TX = 2 * PW, ENT = TY + TX, BINS = PH * PW;
  extern __shared__ __align__(128) unsigned char smem[];
  float* buf = reinterpret_cast<float*>(smem);
  uint2* tab_all = reinterpret_cast<uint2*>(smem + (size_t)c.budget_floats * 4);
  SmemCtl* ctl = reinterpret_cast<SmemCtl*>(smem + (size_t)c.budget_floats * 4 + c.tab_bytes);
  const int tid = threadIdx.x;
  const int Tc = blockDim.x - 32;            // consumer threads; the last warp is the producer
  const int n_cwarps = Tc >> 5;
  if (tid == 0) {
    mbar_init(&ctl->desc_full, 1);
    mbar_init(&ctl->band_full, 1);
    mbar_init(&ctl->band_empty, n_cwarps);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid >= Tc) {
    plane_producer(d, c, w, ctl, buf, tid - Tc, true);
    return;
  }
  const int cw = tid >> 5, lane = tid & 31;
  uint2* slot = tab_all + cw * 16;           // this warp's copy of the current RoI's TY y entries
  const uint4* slot4 = reinterpret_cast<const uint4*>(slot);
  const int xs = min(lane >> 1, TX - 1);     // x sample of this lane
  const bool lane_on = lane < 2 * TX;
  const int t4 = lane & 3, pw = min(lane >> 2, PW - 1);
  const bool odd = lane & 1, up = lane & 2;
  const int o0 = t4 * PW + pw, o1 = (t4 + 4) * PW + pw;
  const bool st0 = lane_on && t4 < PH, st1 = lane_on && t4 + 4 < PH;
  uint32_t phase = 0;
  for (;;) {
    mbar_wait(&ctl->desc_full, phase);
    const ItemSlot it = ctl->item;
    if (it.kind == 2) break;
    if (it.kind == 1) {
      mbar_wait(&ctl->band_full, phase);
      phase ^= 1;
      const RoiGeom g = roi_geom(d, rois, levels, it.fb_roi, c.PH, c.PW, c.sr, c.finest);
      gather_roi_chunk<false>(d, g, it.fb_roi, it.fb_c0, min(32, c.C - it.fb_c0), out, c.PH, c.PW, tid, Tc);
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->band_empty);
      continue;
    }
    // ids two RoIs ahead, tables one RoI ahead: no global latency between jobs
    const int* list = w.list + it.lst;
    int id0 = cw < it.cnt ? list[cw] : -1;
    int id1 = cw + n_cwarps < it.cnt ? list[cw + n_cwarps] : -1;
    uint2 ye = make_uint2(0u, 0u), xe = make_uint2(0u, 0u);
    if (id0 >= 0) {
      if (lane < TY) ye = w.tab[(size_t)id0 * ENT + lane];
      xe = w.tab[(size_t)id0 * ENT + TY + xs];
    }
    mbar_wait(&ctl->band_full, phase);
    phase ^= 1;
    if (!it.bulk) {   // rows not 16-byte aligned: cooperative load by the consumers
      const PlanLevel& v = c.lv[it.lvl];
      const int chan_floats = it.chan_bytes >> 2;
      const float* src0 = d.feat[it.lvl] + (((size_t)it.img * c.C + it.c0) * v.H + it.r0) * v.W;
      for (int i = tid; i < it.ncur * chan_floats; i += Tc) {
        const int j = i / chan_floats;
        buf[i] = __ldg(src0 + (size_t)j * v.H * v.W + (i - j * chan_floats));
      }
      consumer_sync(Tc);
    }
    const int pitch_bytes = it.pitch_bytes;
    for (int k = cw; k < it.cnt; k += n_cwarps) {
      const int n = id0;
      __syncwarp();                       // the previous job's reads of the slot are done
      if (lane < TY) slot[lane] = ye;
      const uint2 xc = xe;
      __syncwarp();
      {
        const int q2 = k + 2 * n_cwarps;
        const int id2 = q2 < it.cnt ? list[q2] : -1;
        if (id1 >= 0) {
          if (lane < TY) ye = w.tab[(size_t)id1 * ENT + lane];
          xe = w.tab[(size_t)id1 * ENT + TY + xs];
        }
        id0 = id1; id1 = id2;
      }
      const float lx = __uint_as_float(xc.y);
      const float wx = lane_on ? (odd ? lx : 1.0f - lx) * 0.25f : 0.0f;
      const char* px = reinterpret_cast<const char*>(buf) + xc.x + (odd ? 4 : 0);
      float* o = out + ((size_t)n * c.C + it.c0) * BINS;
      for (int j = 0; j < it.ncur; ++j) {
        float acc[8];
#pragma unroll
        for (int ph = 0; ph < PH; ++ph) {
          const uint4 e = slot4[ph];      // samples 2ph, 2ph+1: {row offset, l} each
          const char* r0 = px + e.x;
          const char* r1 = px + e.z;
          const float v00 = ldf(r0), v01 = ldf(r0 + pitch_bytes), v10 = ldf(r1), v11 = ldf(r1 + pitch_bytes);
          const float a = fmaf(__uint_as_float(e.y), v01 - v00, v00);
          const float b = fmaf(__uint_as_float(e.w), v11 - v10, v10);
          acc[ph] = (a + b) * wx;
        }
#pragma unroll
        for (int ph = PH; ph < 8; ++ph) acc[ph] = 0.0f;
        // fold the 4 lanes of every bin; lane (pw, t) ends with bins ph = t and ph = t + 4
        float r[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float keep = odd ? acc[2 * q + 1] : acc[2 * q];
          const float send = odd ? acc[2 * q] : acc[2 * q + 1];
          r[q] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
        }
        float s2[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const float keep = up ? r[2 * q + 1] : r[2 * q];
          const float send = up ? r[2 * q] : r[2 * q + 1];
          s2[q] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
        if (st0) o[o0] = s2[0];
        if (st1) o[o1] = s2[1];
        px += it.chan_bytes;
        o += BINS;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&ctl->band_empty);
  }
}

// ------------------------------------- forward, 7x7 / sample_ratio 2, stream kernel ---
// One CTA per SM: 31 consumer warps + one producer warp, TWO band buffers and two table buffers.
// A work unit is (image, level, band, chunk of kS4KC channel groups); the producer lands the band's packed
// RoI tables once per unit (one bulk copy from the band-grouped table) and then streams the unit's channel
// planes through the two band buffers, so the next plane is in flight while the current one is pooled.
// Consumer warps pull (RoI) jobs of the current buffer off a shared-memory counter: no warp waits for
// another one inside a buffer, and a warp that finds the buffer drained moves on to the next.  Per job the
// tap-lane scheme of roi_align_plane_fwd_tap_kernel applies; tables are read straight from the unit's
// shared-memory copy (no per-job global traffic at all).
struct S4Desc {       // 32 bytes: read with two LDS.128
  int kind;            // 0 band, 1 gather fallback, 2 stop | first << 8 | last << 9 | table buffer << 10
  int cnt;             // RoIs in the table buffer            (fallback: RoI id)
  int ncur, c0;        // channels in the band buffer, first  (fallback: -, first channel)
  int chan_bytes, pitch_bytes, pad0, pad1;
};
struct S4Ctl {
  u64 full[kS4Bufs], empty[kS4Bufs], tfull[2], tempty[2];
  S4Desc desc[kS4Bufs];    // 16-byte aligned: an even number of barriers precedes
};
static_assert(sizeof(S4Ctl) <= kS4CtlBytes, "control block");

__device__ __forceinline__ void s4_producer(const FpnDesc& d, const PlanCfg& c, const PlanWs& w, S4Ctl* ctl,
                                            unsigned char* smem, int lane) {
  const int n_fb = w.hdr[1];
  const int fb_chunks = (c.C + 31) / 32;
  const int total = c.n_units + n_fb * fb_chunks;
  const size_t buf_bytes = (size_t)c.budget_floats * 4;
  uint32_t m = 0, u = 0;
  auto publish = [&](const S4Desc& ds, const float* src0, size_t plane_sz) {
    const int b = m % kS4Bufs;
    mbar_wait(&ctl->empty[b], ((m / kS4Bufs) & 1u) ^ 1u);
    if (lane == 0) {
      ctl->desc[b] = ds;
      if ((ds.kind & 0xff) == 0) {
        mbar_arrive_expect_tx(&ctl->full[b], (uint32_t)(ds.ncur * ds.chan_bytes));
        if ((size_t)ds.chan_bytes == plane_sz * 4) {      // whole planes are contiguous: one copy
          bulk_g2s(smem + b * buf_bytes, src0, (uint32_t)(ds.ncur * ds.chan_bytes), &ctl->full[b]);
        } else {
          for (int j = 0; j < ds.ncur; ++j)
            bulk_g2s(smem + b * buf_bytes + (size_t)j * ds.chan_bytes, src0 + (size_t)j * plane_sz,
                     (uint32_t)ds.chan_bytes, &ctl->full[b]);
        }
      } else {
        mbar_arrive(&ctl->full[b]);
      }
    }
    ++m;
  };
  for (;;) {
    int unit = 0;
    if (lane == 0) unit = atomicAdd(&w.hdr[0], 1);
    unit = __shfl_sync(0xffffffffu, unit, 0);
    if (unit >= total) break;
    S4Desc ds;
    ds.kind = 0; ds.cnt = ds.ncur = ds.c0 = ds.chan_bytes = ds.pitch_bytes = ds.pad0 = ds.pad1 = 0;
    if (unit >= c.n_units) {
      const int f = unit - c.n_units;
      ds.kind = 1; ds.cnt = w.fb_list[f / fb_chunks]; ds.c0 = (f % fb_chunks) * 32;
      publish(ds, nullptr, 0);
      continue;
    }
    int l = c.L - 1;                      // unit_base decreases with the level index
    while (l > 0 && unit >= c.lv[l - 1].unit_base) --l;
    const PlanLevel& v = c.lv[l];
    int r = unit - v.unit_base;
    const int chunk = r % v.nchunk; r /= v.nchunk;
    const int band = r % v.nbands, img = r / v.nbands;
    const int bidx = img * c.bands_per_img + v.band_base + band;
    const int cnt = w.cnt[bidx];
    if (cnt == 0) continue;
    const int lst = w.start[bidx];
    const int r0 = band * v.band_rows;
    // whole-plane levels always load whole planes: one contiguous (and 16-byte aligned) copy per group
    const int nrows = v.nbands == 1 ? v.H : min(w.rmax[bidx], v.H - 1) - r0 + 1;
    const size_t plane_sz = (size_t)v.H * v.W;
    const int i0 = chunk * v.kc, i1 = min(v.ncg, i0 + v.kc);
    ds.chan_bytes = nrows * v.W * 4;
    ds.pitch_bytes = v.W * 4;
    for (int rc = 0; rc < cnt; rc += c.s4_max_rois) {
      const int nr = min(c.s4_max_rois, cnt - rc);
      const int tb = u & 1;
      mbar_wait(&ctl->tempty[tb], ((u >> 1) & 1u) ^ 1u);
      if (lane == 0) {
        mbar_arrive_expect_tx(&ctl->tfull[tb], (uint32_t)(nr * c.s4_ent * 8));
        bulk_g2s(smem + kS4Bufs * buf_bytes + (size_t)tb * kS4TabBytes, w.tabg + (size_t)(lst + rc) * c.s4_ent,
                 (uint32_t)(nr * c.s4_ent * 8), &ctl->tfull[tb]);
      }
      ds.cnt = nr;
      for (int i = i0; i < i1; ++i) {
        ds.kind = 0 | ((i == i0) << 8) | ((i == i1 - 1) << 9) | (tb << 10);
        ds.c0 = i * v.cg;
        ds.ncur = min(v.cg, c.C - ds.c0);
        publish(ds, d.feat[l] + (((size_t)img * c.C + ds.c0) * v.H + r0) * v.W, plane_sz);
      }
      ++u;
    }
  }
  S4Desc ds;
  ds.kind = 2; ds.cnt = ds.ncur = ds.c0 = ds.chan_bytes = ds.pitch_bytes = ds.pad0 = ds.pad1 = 0;
  publish(ds, nullptr, 0);
}

template <int PH, int PW>
__global__ void __launch_bounds__((4 * PW <= 32) ? 1024 : 768, 1)     // 14x14 carries 14 accumulators: 85 registers
roi_align_stream_fwd_kernel(const __grid_constant__ FpnDesc d, const __grid_constant__ PlanCfg c, PlanWs w,
                            const float* __restrict__ rois, const int* __restrict__ levels,
                            float* __restrict__ out) {
  // lane = x TAP when the 4*PW taps fit a warp (7x7), else lane = x SAMPLE with both taps loaded by the lane (14x14)
  constexpr bool TAPLANE = 4 * PW <= 32;
  static_assert(TAPLANE ? PH <= 8 : (2 * PW <= 32 && PH % 2 == 0 && PH <= 16), "lane mapping");
  constexpr int TY = 2 * PH, TX = 2 * PW, BINS = PH * PW;
  constexpr int ENT = (TY + TX + 2) & ~1;
  extern __shared__ __align__(128) unsigned char smem[];
  const size_t buf_bytes = (size_t)c.budget_floats * 4;
  S4Ctl* ctl = reinterpret_cast<S4Ctl*>(smem + kS4Bufs * buf_bytes + 2 * kS4TabBytes);
  const int tid = threadIdx.x;
  const int Tc = blockDim.x - 32;
  const int n_cwarps = Tc >> 5;
  if (tid == 0) {
    for (int i = 0; i < kS4Bufs; ++i) {
      mbar_init(&ctl->full[i], 1);
      mbar_init(&ctl->empty[i], n_cwarps);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->tfull[i], 1);
      mbar_init(&ctl->tempty[i], n_cwarps);
    }
    fence_mbar_init();
  }
  __syncthreads();
  if (tid >= Tc) {
    s4_producer(d, c, w, ctl, smem, tid - Tc);
    return;
  }
  const int lane = tid & 31, cw = tid >> 5;
  const int xs = TAPLANE ? min(lane >> 1, TX - 1) : min(lane, TX - 1);
  const bool lane_on = TAPLANE ? lane < 2 * TX : lane < TX;
  const int t4 = lane & 3, pw = TAPLANE ? min(lane >> 2, PW - 1) : min(lane >> 1, PW - 1);
  const bool odd = lane & 1, up = lane & 2;
  const int o0 = t4 * PW + pw, o1 = (t4 + 4) * PW + pw;
  const bool st0 = lane_on && t4 < PH, st1 = lane_on && t4 + 4 < PH;
  const int tap_off = (TAPLANE && odd) ? 4 : 0;
  uint32_t m = 0, u = 0;
  int rot = 0;                   // job rotation: += 11 (mod consumer warps) per message
  for (;;) {
    const int b = m % kS4Bufs;
    mbar_wait(&ctl->full[b], (m / kS4Bufs) & 1u);
    const int4 d0 = reinterpret_cast<const int4*>(&ctl->desc[b])[0];
    const int4 d1 = reinterpret_cast<const int4*>(&ctl->desc[b])[1];
    const int kind = d0.x & 0xff;
    if (kind == 2) break;
    if (kind == 1) {
      const int fb_roi = d0.y, fb_c0 = d0.w;
      const RoiGeom g = roi_geom(d, rois, levels, fb_roi, c.PH, c.PW, c.sr, c.finest);
      gather_roi_chunk<false>(d, g, fb_roi, fb_c0, min(32, c.C - fb_c0), out, c.PH, c.PW, tid, Tc);
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->empty[b]);
      ++m;
      continue;
    }
    const int tb = (d0.x >> 10) & 1, cnt = d0.y, ncur = d0.z, c0 = d0.w;
    const int chan_bytes = d1.x, pitch_bytes = d1.y;
    const bool first = (d0.x >> 8) & 1, last = (d0.x >> 9) & 1;
    if (first) mbar_wait(&ctl->tfull[tb], (u >> 1) & 1u);
    const unsigned char* tabs = smem + kS4Bufs * buf_bytes + (size_t)tb * kS4TabBytes;
    const char* bufb = reinterpret_cast<const char*>(smem + b * buf_bytes) + tap_off;
    // jobs of a buffer are equal-sized: static round-robin, rotated per message so that the warps that get
    // the extra job of a partial round change from buffer to buffer (no counter, no atomics)
    int k0 = cw + rot;
    if (k0 >= n_cwarps) k0 -= n_cwarps;
    rot += 11;
    if (rot >= n_cwarps) rot -= n_cwarps;
    for (int k = k0; k < cnt; k += n_cwarps) {
      const uint2* te = reinterpret_cast<const uint2*>(tabs + (size_t)k * (ENT * 8));
      const uint4* y4 = reinterpret_cast<const uint4*>(te);
      const int n = (int)te[TY + TX].x;
      const uint2 xc = te[TY + xs];
      const float lx = __uint_as_float(xc.y);
      const char* px = bufb + xc.x;
      float* o = out + ((size_t)n * c.C + c0) * BINS;
      if constexpr (TAPLANE) {
        const float wx = lane_on ? (odd ? lx : 1.0f - lx) * 0.25f : 0.0f;
        for (int j = 0; j < ncur; ++j) {
          float acc[8];
#pragma unroll
          for (int ph = 0; ph < PH; ++ph) {
            const uint4 e = y4[ph];         // samples 2ph, 2ph+1: {row offset, l} each
            const char* r0 = px + e.x;
            const char* r1 = px + e.z;
            const float v00 = ldf(r0), v01 = ldf(r0 + pitch_bytes), v10 = ldf(r1), v11 = ldf(r1 + pitch_bytes);
            const float a = fmaf(__uint_as_float(e.y), v01 - v00, v00);
            const float bq = fmaf(__uint_as_float(e.w), v11 - v10, v10);
            acc[ph] = (a + bq) * wx;
          }
#pragma unroll
          for (int ph = PH; ph < 8; ++ph) acc[ph] = 0.0f;
          float r[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float keep = odd ? acc[2 * q + 1] : acc[2 * q];
            const float send = odd ? acc[2 * q] : acc[2 * q + 1];
            r[q] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
          }
          float s2[2];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const float keep = up ? r[2 * q + 1] : r[2 * q];
            const float send = up ? r[2 * q] : r[2 * q + 1];
            s2[q] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
          }
          if (st0) o[o0] = s2[0];
          if (st1) o[o1] = s2[1];
          px += chan_bytes;
          o += BINS;
        }
      } else {
        // lane = x sample: the lane loads both taps of its sample ([r], [r+4]: one address, one immediate) on
        // both rows, interpolates in x then y; the two lanes of a bin swap half of their PH sums
        const float* ob = o + pw;
        for (int j = 0; j < ncur; ++j) {
          float acc[PH];
#pragma unroll
          for (int ph = 0; ph < PH; ++ph) {
            // volatile load: re-read the row entries per channel instead of parking 56 registers across the loop
            uint4 e;
            asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(e.x), "=r"(e.y), "=r"(e.z), "=r"(e.w)
                         : "r"(smem_u32(y4 + ph)));
            const char* r0 = px + e.x;
            const char* r1 = px + e.z;
            const float a00 = ldf(r0), a01 = ldf(r0 + 4), a10 = ldf(r0 + pitch_bytes), a11 = ldf(r0 + pitch_bytes + 4);
            const float b00 = ldf(r1), b01 = ldf(r1 + 4), b10 = ldf(r1 + pitch_bytes), b11 = ldf(r1 + pitch_bytes + 4);
            const float at = fmaf(lx, a01 - a00, a00), ab = fmaf(lx, a11 - a10, a10);
            const float bt = fmaf(lx, b01 - b00, b00), bb = fmaf(lx, b11 - b10, b10);
            acc[ph] = fmaf(__uint_as_float(e.y), ab - at, at) + fmaf(__uint_as_float(e.w), bb - bt, bt);
          }
#pragma unroll
          for (int q = 0; q < PH / 2; ++q) {
            const float keep = odd ? acc[2 * q + 1] : acc[2 * q];
            const float send = odd ? acc[2 * q] : acc[2 * q + 1];
            const float v = (keep + __shfl_xor_sync(0xffffffffu, send, 1)) * 0.25f;
            if (lane_on) const_cast<float*>(ob)[(2 * q + (odd ? 1 : 0)) * PW] = v;
          }
          px += chan_bytes;
          ob += BINS;
        }
      }
    }
    __syncwarp();
    if (lane == 0) {
      if (last) mbar_arrive(&ctl->tempty[tb]);
      mbar_arrive(&ctl->empty[b]);
    }
    if (last) ++u;
    ++m;
  }
}

// --------------------------------------------------------------- backward --------
// L2 RED.ADD throughput on B200 is bound per 32-byte SECTOR touched by a warp-level RED instruction
// (~200 G sector-ops/s, profiles/microbench), not per lane, and shared-memory fp32 atomics are CAS
// loops that are no faster (a band-resident CAS variant measured 1.87 ms on BASELINE config 3 against
// 1.80 ms for per-tap REDs).  So the backward minimises sector-ops instead:
//   * lane gl of a 16/32-lane group owns x-sample gl of one (RoI, channel): the lanes of one RED
//     instruction hit a few adjacent sectors of ONE gradient row;
//   * the high tap of lane gl and the low tap of lane gl+1 usually fall on the same column: it is handed
//     over with one shuffle and added once;
//   * the high row of sample row sy and the low row of sy+1 usually coincide: it is carried in
//     registers and written once.
// Both merges are optimisations only - every write is still an atomic RED, so overlapping RoIs,
// coincident samples and clamped borders stay exact.  One CTA per (RoI, 32-channel chunk).
constexpr int kRowsBwdThreads = 256;

template <int SR>
__global__ void __launch_bounds__(kRowsBwdThreads)
roi_align_rows_bwd_kernel(const __grid_constant__ FpnDesc d, const __grid_constant__ PlanCfg c, PlanWs w,
                          const float* __restrict__ rois, const int* __restrict__ levels,
                          float* __restrict__ gout, int cchunk) {
  __shared__ uint2 te[128];
  const int n = blockIdx.x;
  const int c0 = blockIdx.y * cchunk;
  const int cc = min(cchunk, c.C - c0);
  const int tid = threadIdx.x;
  const RoiGeom g = roi_geom(d, rois, levels, n, c.PH, c.PW, c.sr, c.finest);
  if (w.meta[n] < 0) {   // bad index / samples outside the image: generic per-tap path
    gather_roi_chunk<true>(d, g, n, c0, cc, gout, c.PH, c.PW, tid, kRowsBwdThreads);
    return;
  }
  const int ent = c.ty + c.tx;
  const int sr = SR > 0 ? SR : c.sr;
  const int bins = c.PH * c.PW;
  for (int t = tid; t < ent; t += kRowsBwdThreads) te[t] = w.tab[(size_t)n * ent + t];
  __syncthreads();
  // ---- dense-column mode -------------------------------------------------------------------------
  // When the RoI's footprint is at most 32 columns wide and every column is fed by samples of at most
  // two neighbouring bins (bin width >= 1 px), lane i of a group owns footprint column xmin+i: the
  // x-direction of the adjoint collapses to  e[ph][col] = (A_col*g[ph][p0_col] + B_col*g[ph][p0_col+1])/count
  // with per-column constants computed once per CTA, and every gradient row is ONE contiguous RED per
  // group - the minimum number of L2 sectors.  Other RoIs (tiny or very wide) use the per-sample path below.
  {
    const int xmin = (int)(te[c.ty].x >> 2);
    const int Wf = (int)(te[c.ty + c.tx - 1].x >> 2) + 2 - xmin;
    const int gsz = Wf <= 16 ? 16 : 32;
    const int col = tid % gsz;
    float A = 0.0f, B = 0.0f;
    int p0 = -1, pmax = -1;
    if (Wf <= 32 && col < Wf) {
      for (int sx = 0; sx < c.tx; ++sx) {
        const uint2 e = te[c.ty + sx];
        const int xs = (int)(e.x >> 2) - xmin;
        const float l = __uint_as_float(e.y);
        float wgt = 0.0f;
        bool hit = false;
        if (xs == col) { wgt = 1.0f - l; hit = true; }
        if (xs + 1 == col) { wgt = l; hit = true; }
        if (hit) {
          const int pb = sx / sr;
          if (p0 < 0) p0 = pb;
          if (pb == p0) A += wgt; else B += wgt;
          pmax = pb;
        }
      }
    }
    const int bad = (Wf > 32) || (p0 >= 0 && pmax > p0 + 1);
    if (!__syncthreads_or(bad)) {
      const int grp = tid / gsz, ngroups = kRowsBwdThreads / gsz;
      const bool lane_on = col < Wf && p0 >= 0;
      const float inv_count = 1.0f / (float)(sr * sr);
      const int W = g.W;
      const size_t plane_sz = (size_t)g.H * W;
      const bool has_b = lane_on && (p0 + 1 < c.PW) && B != 0.0f;
      const float a_s = A * inv_count, b_s = B * inv_count;
      for (int cj = grp; cj < cc; cj += ngroups) {
        float* plane = g.plane0 + (size_t)(c0 + cj) * plane_sz + xmin + col;
        const float* gj = gout + ((size_t)n * c.C + c0 + cj) * bins + (lane_on ? p0 : 0);
        float carry = 0.0f;
        int carry_row = -1;
        for (int ph = 0; ph < c.PH; ++ph) {
          float e = 0.0f;
          if (lane_on) {
            e = a_s * __ldg(gj + ph * c.PW);
            if (has_b) e += b_s * __ldg(gj + ph * c.PW + 1);
          }
          for (int iy = 0; iy < sr; ++iy) {
            const uint2 ey = te[ph * sr + iy];
            const int row = (int)ey.x;
            const float ly = __uint_as_float(ey.y), hy = 1.0f - ly;
            float a = e * hy;
            if (row == carry_row) {
              a += carry;
            } else if (carry_row >= 0 && lane_on) {
              atomicAdd(plane + (size_t)carry_row * W, carry);
            }
            if (lane_on) atomicAdd(plane + (size_t)row * W, a);
            carry_row = row + 1;
            carry = e * ly;
          }
        }
        if (carry_row >= 0 && lane_on) atomicAdd(plane + (size_t)carry_row * W, carry);
      }
      return;
    }
  }
  // ---- per-sample mode ---------------------------------------------------------------------------
  const int grp = tid / c.group, gl = tid % c.group, ngroups = kRowsBwdThreads / c.group;
  const bool lane_on = gl < c.tx;
  const uint2 ex = te[c.ty + (lane_on ? gl : 0)];
  const int xo = (int)(ex.x >> 2);
  const float lx = __uint_as_float(ex.y), hx = 1.0f - lx;
  // column hand-over between neighbouring lanes of the group
  const int xo_prev = __shfl_up_sync(0xffffffffu, xo, 1), xo_next = __shfl_down_sync(0xffffffffu, xo, 1);
  const bool take_prev = lane_on && gl > 0 && xo == xo_prev + 1;
  const bool give_next = lane_on && gl + 1 < c.tx && xo_next == xo + 1;
  const float inv_count = 1.0f / (float)(sr * sr);
  const int W = g.W;
  const size_t plane_sz = (size_t)g.H * W;
  const int rounds = (cc + ngroups - 1) / ngroups;
  for (int k = 0; k < rounds; ++k) {
    const int cj = k * ngroups + grp;
    const bool live = cj < cc && lane_on;
    float* plane = g.plane0 + (size_t)(c0 + (cj < cc ? cj : 0)) * plane_sz + xo;
    const float* gj = gout + ((size_t)n * c.C + c0 + (cj < cc ? cj : 0)) * bins + (lane_on ? gl / sr : 0);
    float carry_lo = 0.0f, carry_hi = 0.0f;
    int carry_row = -1;
    // one merged row write: hand the high tap to the next lane when it owns that column
    auto flush = [&](int row, float lo, float hi) {
      const float recv = __shfl_up_sync(0xffffffffu, hi, 1);
      if (take_prev) lo += recv;
      if (live) {
        float* r = plane + (size_t)row * W;
        atomicAdd(r, lo);
        if (!give_next) atomicAdd(r + 1, hi);
      }
    };
    float g_next = __ldg(gj);
    for (int ph = 0; ph < c.PH; ++ph) {
      const float gv = g_next * inv_count;
      if (ph + 1 < c.PH) g_next = __ldg(gj + (ph + 1) * c.PW);
      const float ghx = gv * hx, glx = gv * lx;
#pragma unroll
      for (int iy = 0; iy < (SR > 0 ? SR : 1); ++iy) {
        const uint2 ey = te[ph * sr + iy];
        const int row = (int)ey.x;
        const float ly = __uint_as_float(ey.y), hy = 1.0f - ly;
        float a_lo = ghx * hy, a_hi = glx * hy;
        if (row == carry_row) {          // group-uniform: the carried high row is this low row
          a_lo += carry_lo; a_hi += carry_hi;
        } else if (carry_row >= 0) {
          flush(carry_row, carry_lo, carry_hi);
        }
        flush(row, a_lo, a_hi);
        carry_row = row + 1; carry_lo = ghx * ly; carry_hi = glx * ly;
      }
      if (SR == 0) {
        for (int iy = 1; iy < sr; ++iy) {
          const uint2 ey = te[ph * sr + iy];
          const int row = (int)ey.x;
          const float ly = __uint_as_float(ey.y), hy = 1.0f - ly;
          float a_lo = ghx * hy, a_hi = glx * hy;
          if (row == carry_row) {
            a_lo += carry_lo; a_hi += carry_hi;
          } else if (carry_row >= 0) {
            flush(carry_row, carry_lo, carry_hi);
          }
          flush(row, a_lo, a_hi);
          carry_row = row + 1; carry_lo = ghx * ly; carry_hi = glx * ly;
        }
      }
    }
    if (carry_row >= 0) flush(carry_row, carry_lo, carry_hi);
  }
}

}  // namespace mxd

namespace mxd {

size_t plane_workspace_bytes(int R, int N, int L, const int* Hs, const int* Ws, int C, int PH, int PW, int sr) {
  PlanCfg c;
  size_t need = 256;
  if (make_cfg(N, C, L, Hs, Ws, PH, PW, sr, 56.0f, &c)) need = carve_plan(nullptr, R, c.NB, c.ty + c.tx).bytes;
  if (make_cfg(N, C, L, Hs, Ws, PH, PW, sr, 56.0f, &c, true)) {
    const size_t b = carve_plan(nullptr, R, c.NB, c.ty + c.tx).bytes;
    if (b > need) need = b;
  }
  return need;
}

static int num_sms() {      // per call: the current device may differ between calls
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms > 0 ? sms : 148;
}

static int run_planner(const FpnDesc& d, const PlanCfg& c, const PlanWs& w, const float* rois, const int* levels,
                       int R, int bwd, cudaStream_t st) {
  // zero hdr, cnt, rmax, start, cursor in one go (they are contiguous)
  const size_t zbytes = (size_t)((char*)w.meta - (char*)w.hdr);
  MXD_CUDA_OK(cudaMemsetAsync(w.hdr, 0, zbytes, st));
  if (R == 0) return MXD_OK;
  plan_rois_kernel<<<(R * 32 + 255) / 256, 256, 0, st>>>(d, c, w, rois, levels, R, bwd);
  MXD_POST_LAUNCH("roi_align_plan_rois");
  plan_group_kernel<<<1, 1024, 0, st>>>(c, w, R, bwd);
  MXD_POST_LAUNCH("roi_align_plan_group");
  return MXD_OK;
}

int plane_forward(const FpnDesc& d, const float* rois, const int* levels, float* out, int R, int PH, int PW, int sr,
                  float finest, void* ws, size_t ws_bytes, cudaStream_t st, int* handled) {
  *handled = 0;
  PlanCfg c;
  if (R == 0 || d.C == 0) return MXD_OK;
  bool stream = make_cfg(d.N, d.C, d.num_levels, d.H, d.W, PH, PW, sr, finest, &c, true);
  for (int l = 0; stream && l < d.num_levels; ++l)
    if ((reinterpret_cast<uintptr_t>(d.feat[l]) & 15) != 0) stream = false;     // TMA source alignment
  if (!stream && !make_cfg(d.N, d.C, d.num_levels, d.H, d.W, PH, PW, sr, finest, &c)) return MXD_OK;
  PlanWs w = carve_plan(ws, R, c.NB, c.ty + c.tx);
  MXD_REQUIRE(ws_bytes >= w.bytes, MXD_EWORKSPACE, "roi_align workspace %zu < %zu bytes", ws_bytes, w.bytes);
  MXD_REQUIRE(((uintptr_t)ws & 255) == 0, MXD_EINVAL, "workspace must be 256-byte aligned");
  int rc;
  if ((rc = run_planner(d, c, w, rois, levels, R, 0, st))) return rc;
  if (stream) {
    plan_pack_kernel<<<(R * c.s4_ent + 255) / 256, 256, 0, st>>>(c, w, R);
    MXD_POST_LAUNCH("roi_align_plan_pack");
    static unsigned long long seen4 = 0;
    DeviceOnce once_seen4(&seen4);
  if (once_seen4.first()) {
      MXD_CUDA_OK(cudaFuncSetAttribute(roi_align_stream_fwd_kernel<7, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kSmemLimit));
      MXD_CUDA_OK(cudaFuncSetAttribute(roi_align_stream_fwd_kernel<14, 14>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kSmemLimit));
    }
    if (PH == 7)
      roi_align_stream_fwd_kernel<7, 7><<<num_sms(), c.threads, c.smem_bytes, st>>>(d, c, w, rois, levels, out);
    else
      roi_align_stream_fwd_kernel<14, 14><<<num_sms(), c.threads, c.smem_bytes, st>>>(d, c, w, rois, levels, out);
    MXD_POST_LAUNCH("roi_align_stream_fwd");
    *handled = 1;
    return MXD_OK;
  }
  const bool tap = sr == 2 && PH == 7 && PW == 7 && c.tab_bytes >= ((c.threads - 32) >> 5) * 128;
  auto kern = tap ? roi_align_plane_fwd_tap_kernel<7, 7>
                  : (sr == 2) ? roi_align_plane_fwd_kernel<2> : roi_align_plane_fwd_kernel<0>;
  static unsigned long long seen[3] = {0, 0, 0};
  const int ki = tap ? 2 : sr == 2 ? 0 : 1;
  DeviceOnce once_k(&seen[ki]);
  if (once_k.first())
    MXD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
  kern<<<num_sms() * c.ctas_per_sm, c.threads, c.smem_bytes, st>>>(d, c, w, rois, levels, out);
  MXD_POST_LAUNCH("roi_align_plane_fwd");
  *handled = 1;
  return MXD_OK;
}

int plane_backward(const FpnDesc& d, const float* rois, const int* levels, const float* gout, int R, int PH, int PW,
                   int sr, float finest, int accumulate, void* ws, size_t ws_bytes, cudaStream_t st, int* handled) {
  *handled = 0;
  PlanCfg c;
  if (R == 0 || d.C == 0) return MXD_OK;
  if (!make_cfg(d.N, d.C, d.num_levels, d.H, d.W, PH, PW, sr, finest, &c)) return MXD_OK;
  PlanWs w = carve_plan(ws, R, c.NB, c.ty + c.tx);
  MXD_REQUIRE(ws_bytes >= w.bytes, MXD_EWORKSPACE, "roi_align workspace %zu < %zu bytes", ws_bytes, w.bytes);
  MXD_REQUIRE(((uintptr_t)ws & 255) == 0, MXD_EINVAL, "workspace must be 256-byte aligned");
  if (!accumulate)
    for (int l = 0; l < c.L; ++l)
      MXD_CUDA_OK(cudaMemsetAsync(d.feat[l], 0, sizeof(float) * (size_t)d.N * d.C * d.H[l] * d.W[l], st));
  plan_rois_kernel<<<(R * 32 + 255) / 256, 256, 0, st>>>(d, c, w, rois, levels, R, 1);
  MXD_POST_LAUNCH("roi_align_plan_rois");
  const int ngroups = kRowsBwdThreads / c.group;
  int cchunk = 2 * ngroups;                       // two channels per lane group per CTA
  while ((d.C + cchunk - 1) / cchunk > 65535) cchunk *= 2;
  dim3 grid(R, (d.C + cchunk - 1) / cchunk);
  if (sr == 2)
    roi_align_rows_bwd_kernel<2><<<grid, kRowsBwdThreads, 0, st>>>(d, c, w, rois, levels, const_cast<float*>(gout), cchunk);
  else
    roi_align_rows_bwd_kernel<0><<<grid, kRowsBwdThreads, 0, st>>>(d, c, w, rois, levels, const_cast<float*>(gout), cchunk);
  MXD_POST_LAUNCH("roi_align_rows_bwd");
  *handled = 1;
  return MXD_OK;
}

}  // namespace mxd
