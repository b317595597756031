// Row-ring RoIAlign forward for B200 (Spec A, G): 7x7 bins, sample_ratio 2.
//
// Contract as roi_align.cu (mx.nd.contrib.ROIAlign, mxdetection/ops, /root/reference/README.md:24;
// SingleLevelRoI.forward, /root/reference/README.md:32).  Design:
//
//  * One persistent CTA per SM.  A work unit is (image, level, group of `cg` channels).  The unit's channel planes
//    STREAM through shared memory exactly once, top to bottom, in chunks of CH rows: chunk k of channel j is ONE
//    2-D tensor-map TMA copy (cp.async.bulk.tensor.2d, SASS UTMALDG; box = CH rows x a fixed number of columns,
//    columns past W zero-filled) landing at ring rows (k mod NS)*CH.. of channel j's region.  The box width makes
//    the shared-memory ROW PITCH A COMPILE-TIME CONSTANT of the level's class (352 / 704 / 1408 / 2048 bytes with
//    8 / 4 / 2 / 1 channels per unit), and channel j's region starts at j * (arena / cg): every tap address of the
//    channel loop is [register + immediate].  The ring row of map row y is y mod (NS*CH) for every unit, so the
//    planner writes final shared-memory offsets into the sample tables; the row after the last ring row is a copy
//    of ring row 0, which keeps "high tap = low tap + pitch" true across the wrap.
//  * RoIs are grouped by the chunk holding their first tap row and processed in that order, so a chunk is dead
//    as soon as every consumer warp has moved past its group: no band overlap, every map byte crosses L2 -> SM
//    once (the two-buffer band kernel it replaces re-read 40 % of every band and carried ONE channel per message on
//    the two fine FPN levels; here they carry 2 and 4, the coarse ones 8).
//  * A consumer warp owns a (RoI, all cg channels) job: it expands the RoI's packed table ONCE into registers -
//    14 row offsets and 28 tap weights, already multiplied by the lane's x weight - and then runs the channel
//    loop as 28 LDS [R + immediate] + 28 FFMA + a fold (6 shuffles) + 2 stores per (RoI, channel); lane = x tap
//    (sample l>>1, low/high tap l&1), so one LDS reads the 28 taps of a sample row without bank conflicts.
//    (Holding the bin rows in lane-PERMUTED slots makes the fold select-free but lets one LDS read four different
//    rows: 26 M bank-conflict wavefronts, 0.39 ms - rejected.)
//  * Small levels whose rows cannot be bulk-copied one by one (W*4 not a multiple of 16) run in "tall" mode:
//    cg whole planes are one contiguous copy into one of two slots; same consumer code, same barriers.
//  * Synchronisation: full[q & 15] transaction barriers (fills are numbered absolutely: barrier q & 15, parity
//    (q >> 4) & 1), one progress word per consumer warp (the global number of the job it is working on, published
//    with st.release) that the producer polls to learn when a ring slot's rows are dead, and a small ring of unit
//    descriptors.  Jobs are dealt round-robin by global job number across units.  No CTA barrier, no atomics.
//  * RoIs that need more rows than the ring holds, sample outside the image or carry a bad batch index take the
//    generic gather afterwards (same kernel, extra units).
#include <cuda.h>
#include <stdlib.h>
#include "roi_align.cuh"
#include "ptx.cuh"

namespace mxd {

constexpr int kRgSmem = 227 * 1024;
constexpr int kRgThreads = 640;                 // 19 consumer warps + 1 producer warp; 96 registers per thread (768 / 1024 threads spill and are slower)
constexpr int kRgWarps = kRgThreads / 32 - 1;
constexpr int kRgSlotsMax = 16;                 // ring slots (mbarriers)
constexpr int kRgDesc = 4;                      // unit-descriptor ring
constexpr int kRgCtlBytes = 1792;
constexpr int kRgTabSlot = 256;                 // per-warp copy of the current RoI's table
constexpr int kRgArena = (kRgSmem - kRgCtlBytes - kRgWarps * kRgTabSlot) & ~127;
constexpr int kRgCH = 8;                        // rows per chunk in ring mode
constexpr int kRgEnt = 30;                      // table entries per RoI: 14 y + 14 x + {id, kf | ce << 16} + pad
constexpr int kRgMaxNk = 128;                   // chunks per plane (the producer stages the group table in shared memory)
// ring mode: channel j of a unit lives at j * (arena / cg) - a compile-time constant per cg, so the channel loop is
// fully unrolled with the channel offset as the IMMEDIATE of every LDS (no address arithmetic per channel)
__host__ __device__ constexpr int rg_region(int cg) { return (kRgArena / cg) & ~127; }
// level classes k = 0..3: channels per unit and shared-memory row pitch in bytes (= TMA box width; <= 256 8-byte elements)
constexpr int kRgClasses = 4;
__host__ __device__ constexpr int rg_cls_cg(int k) { return 8 >> k; }
__host__ __device__ constexpr int rg_cls_pitch(int k) { return k < 3 ? (352 << k) : 2048; }

struct RgLevel {
  int H, W;
  int mode;          // 1 ring (tensor-map rows, class cls), 2 tall (whole planes, bulk copy)
  int cls, pitch;    // ring class; shared-memory row pitch in bytes (tall: W * 4)
  int cg, ncc;       // channels per unit, channel chunks per image
  int NS;            // slots
  int nk;            // chunks (= RoI groups) per plane: ceil(H / CH) in ring mode, 1 in tall mode
  int region;        // bytes between channels inside the arena
  int slot_bytes;    // tall mode: bytes of one slot
  int group_base;    // first group of this level inside one image's group table
  int unit_base;     // first unit id of this level
};

struct RgCfg {
  RgLevel lv[MXD_MAX_LEVELS];
  int L, N, C;
  int groups_per_img, NB, n_units;
  float finest;
};

struct RgMaps {      // one 2-D tensor map per ring level: dims (W/2 x 8-byte elements, N*C*H rows), box (pitch/8, CH)
  CUtensorMap m[MXD_MAX_LEVELS];
};

struct RgWs {
  int* hdr;      // [0] unit counter, [1] fallback count
  int* cnt;      // [NB] RoIs per group
  int* rmax;     // [NB] last tap row over the group's RoIs
  int* start;    // [NB] exclusive prefix of cnt
  int* meta;     // [R] group index or -1
  int* rank;     // [R] position inside the group
  int* fb_list;  // [R]
  uint2* tab;    // [R][kRgEnt]
  uint2* tabg;   // [R][kRgEnt] in group order
  size_t bytes;
};

static RgWs rg_carve(void* base, int R, int NB) {
  RgWs w;
  size_t off = 0;
  const size_t r1 = (size_t)(R > 0 ? R : 1);
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return (char*)base + o; };
  w.hdr = (int*)take(sizeof(int) * 64);
  w.cnt = (int*)take(sizeof(int) * (size_t)NB);
  w.rmax = (int*)take(sizeof(int) * (size_t)NB);
  w.start = (int*)take(sizeof(int) * (size_t)NB);
  w.meta = (int*)take(sizeof(int) * r1);
  w.rank = (int*)take(sizeof(int) * r1);
  w.fb_list = (int*)take(sizeof(int) * r1);
  w.tab = (uint2*)take(sizeof(uint2) * r1 * kRgEnt);
  w.tabg = (uint2*)take(sizeof(uint2) * r1 * kRgEnt);
  w.bytes = off;
  return w;
}

// Host: per level, ring mode in the narrowest class whose row pitch holds a map row, else tall mode.
static bool rg_make_cfg(const FpnDesc& d, int PH, int PW, int sr, float finest, RgCfg* c, bool have_tma = true) {
  if (sr != 2 || PH != 7 || PW != 7 || d.C <= 0 || d.N <= 0) return false;
  memset(c, 0, sizeof(*c));
  c->L = d.num_levels; c->N = d.N; c->C = d.C; c->finest = finest;
  int gbase = 0;
  for (int l = 0; l < c->L; ++l) {
    RgLevel& v = c->lv[l];
    v.H = d.H[l]; v.W = d.W[l];
    if (v.H < 2 || v.W < 2) return false;
    const size_t row = (size_t)v.W * 4, plane = row * v.H;
    const bool aligned = (reinterpret_cast<uintptr_t>(d.feat[l]) & 15) == 0;
    v.mode = 0;
    if (aligned && (row & 15) == 0 && have_tma) {
      const int nk = (v.H + kRgCH - 1) / kRgCH;
      int k = 0;
      while (k < kRgClasses && (size_t)rg_cls_pitch(k) < row) ++k;
      if (k < kRgClasses && nk <= kRgMaxNk) {
        const int cg = rg_cls_cg(k), pitch = rg_cls_pitch(k);
        const int rows = rg_region(cg) / pitch - 1;                                // one extra row: the wrap copy
        int ns = std::min(rows / kRgCH, kRgSlotsMax);
        if (ns > nk) ns = nk;
        if (ns >= 2 || nk == 1) {
          v.mode = 1; v.cls = k; v.pitch = pitch; v.cg = cg; v.NS = ns; v.nk = nk;
          v.region = rg_region(cg);
        }
      }
    }
    if (!v.mode && aligned) {   // tall mode: cg whole planes per copy, two slots
      int al = 1;
      while ((plane * al) & 15) al <<= 1;
      int cg = 8;
      while (cg > al && (cg > d.C || align_up((size_t)cg * plane, 128) * 2 > (size_t)kRgArena)) cg >>= 1;
      if (cg >= al && d.C % al == 0 && cg <= d.C && align_up((size_t)cg * plane, 128) * 2 <= (size_t)kRgArena) {
        v.mode = 2; v.cg = cg; v.NS = 2; v.nk = 1;
        v.region = (int)plane; v.slot_bytes = (int)align_up((size_t)cg * plane, 128); v.pitch = (int)row;
      }
    }
    if (!v.mode) return false;
    v.ncc = (d.C + v.cg - 1) / v.cg;
    v.group_base = gbase; gbase += v.nk;
  }
  for (int l = c->L - 1; l >= 0; --l) {      // unit ids: coarsest level first (its units carry the most RoIs)
    c->lv[l].unit_base = c->n_units;
    c->n_units += c->N * c->lv[l].ncc;
  }
  c->groups_per_img = gbase;
  c->NB = c->N * gbase;
  return true;
}

// ------------------------------------------------------------------- planner ------
// One warp per RoI: Spec A sample tables with FINAL shared-memory offsets, group = chunk of the first tap row.
__global__ void __launch_bounds__(256) rg_plan_rois_kernel(FpnDesc d, RgCfg c, RgWs w, const float* __restrict__ rois,
                                                            const int* __restrict__ levels, int R) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  pdl_wait();                       // the workspace memset (and whatever produced rois) is complete
  pdl_launch_dependents();
  if (n >= R) return;
  const RoiGeom g = roi_geom(d, rois, levels, n, 7, 7, 2, c.finest);
  const RgLevel& v = c.lv[g.ok ? g.lvl : 0];
  bool ok = g.ok;
  AxisTap ya, xa;
  ya.lo = ya.hi = xa.lo = xa.hi = 0; ya.l = xa.l = 0.0f; ya.valid = xa.valid = 1;
  int rfirst = 0x7fffffff, rlast = -1;
  if (lane < 14) {
    ya = axis_tap(g.rsh, g.bh, 2, lane >> 1, lane & 1, g.H, 1);
    xa = axis_tap(g.rsw, g.bw, 2, lane >> 1, lane & 1, g.W, 1);
    rfirst = ya.hi == ya.lo ? ya.lo - 1 : ya.lo;
    rlast = ya.hi;
  }
  const bool valid = ya.valid && xa.valid;
  rfirst = __reduce_min_sync(0xffffffffu, rfirst);
  rlast = __reduce_max_sync(0xffffffffu, rlast);
  ok = ok && __all_sync(0xffffffffu, valid);
  int kf = 0, ce = 0;
  if (ok && v.mode == 1) {
    kf = rfirst / kRgCH; ce = rlast / kRgCH;
    if (v.NS < v.nk && ce - kf + 1 > v.NS - 1) ok = false;      // the footprint must fit the ring with a slot to spare
  }
  if (!ok) {
    if (lane == 0) {
      w.meta[n] = -1;
      w.fb_list[atomicAdd(&w.hdr[1], 1)] = n;
    }
    return;
  }
  const int bidx = g.b * c.groups_per_img + v.group_base + kf;
  uint2* tab = w.tab + (size_t)n * kRgEnt;
  // Entries are in "unclamped" form (see roi_align_plane.cu): low tap t, high tap t+1, weight l of the high tap; a
  // sample clamped at the border (lo == hi == size-1, l == 0) becomes (size-2, l = 1) - the same value.
  if (lane < 14) {
    const bool cy = ya.hi == ya.lo;
    const int lo = cy ? ya.lo - 1 : ya.lo;
    const int rr = v.mode == 1 ? lo % (v.NS * kRgCH) : lo;
    tab[lane] = make_uint2((unsigned)(rr * v.pitch), __float_as_uint(cy ? 1.0f : ya.l));
    const bool cx = xa.hi == xa.lo;
    tab[14 + lane] = make_uint2((unsigned)((cx ? xa.lo - 1 : xa.lo) * 4), __float_as_uint(cx ? 1.0f : xa.l));
  }
  if (lane == 14) tab[28] = make_uint2((unsigned)n, (unsigned)kf | ((unsigned)ce << 16));
  if (lane == 15) tab[29] = make_uint2(0u, 0u);
  if (lane == 0) {
    w.meta[n] = bidx;
    w.rank[n] = atomicAdd(&w.cnt[bidx], 1);
    atomicMax(&w.rmax[bidx], rlast);
  }
}

// Every CTA scans the group counts into shared memory (a few hundred entries), then each of its warps moves one
// RoI's 240-byte table to its place in group order; CTA 0 also publishes the prefix for the main kernel.
constexpr int kRgMaxGroups = 8192;
__global__ void __launch_bounds__(256) rg_plan_pack_kernel(RgCfg c, RgWs w, int R) {
  __shared__ int s_start[kRgMaxGroups];
  __shared__ int s_part[256];
  const int tid = threadIdx.x;
  pdl_wait();                       // the RoI plans are complete
  pdl_launch_dependents();
  const int per = (c.NB + 255) / 256;
  int sum = 0;
  for (int i = tid * per; i < min(c.NB, (tid + 1) * per); ++i) sum += w.cnt[i];
  s_part[tid] = sum;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {
    const int v = tid >= o ? s_part[tid - o] : 0;
    __syncthreads();
    s_part[tid] += v;
    __syncthreads();
  }
  int run = s_part[tid] - sum;
  for (int i = tid * per; i < min(c.NB, (tid + 1) * per); ++i) {
    s_start[i] = run;
    if (blockIdx.x == 0) w.start[i] = run;
    run += w.cnt[i];
  }
  __syncthreads();
  const int n = blockIdx.x * 8 + (tid >> 5), lane = tid & 31;
  if (n >= R) return;
  const int m = w.meta[n];
  if (m < 0) return;
  const int pos = s_start[m] + w.rank[n];
  if (lane < kRgEnt) w.tabg[(size_t)pos * kRgEnt + lane] = w.tab[(size_t)n * kRgEnt + lane];
}

// --------------------------------------------------------------- main kernel ------
struct RgDesc {          // 64 bytes
  int kind;              // 0 ring/tall unit, 1 gather fallback, 2 stop
  int ncur, c0, jobs;
  int tstart, jbase, ka, kb;       // first job in tabg; global job number of job 0; chunk range (absolute)
  int qoff, chan_stride, pitch, bufoff;   // fill number of chunk k = qoff + k
  int fb_roi, fb_c0, kbase, mode_cls;     // mode | class << 8
};
// Synchronisation state.  full[q & 15] is the transaction barrier of fill number q (fills are numbered over the whole
// life of the CTA, so barrier and phase parity follow from q alone and a warp only ever waits for chunks it reads).
// A chunk is dead when every job that starts at or before it is done: consumer warp w publishes prog[w], the global
// number of its next job (jobs q = w mod n_warps, in order), and the producer refills a ring slot once the minimum
// over the warps has passed the slot's threshold - one shared-memory store per JOB instead of a barrier round trip per
// (warp, chunk), which cost a third of the instructions of the first version of this kernel.
struct RgCtl {
  u64 full[kRgSlotsMax], dfull[kRgDesc], dempty[kRgDesc];
  RgDesc desc[kRgDesc];
  int prog[32];
  int thresh[kRgSlotsMax];   // producer: job number that frees the ring slot's current content
  int gneed[kRgMaxNk];       // producer: last chunk needed by the groups <= k of the current unit (-1: none)
  int gcum[kRgMaxNk];        // producer: jobs in the groups <= k of the current unit
};
static_assert(sizeof(RgCtl) <= kRgCtlBytes, "control block");

// Optional phase timers (-DMXD_RING_PROF): cycles blocked on the ring, accumulated into hdr[16..]
#ifdef MXD_RING_PROF
#define RG_T0() const long long _t0 = clock64()
#define RG_T1(acc) acc += clock64() - _t0
#else
#define RG_T0()
#define RG_T1(acc)
#endif

__device__ __forceinline__ void tma_2d_g2s(void* dst, const CUtensorMap* map, int x, int y, u64* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ int lds_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void sts_release(int* p, int v) {
  asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
// parity wait on the barrier at shared address `bar` (spin inside the asm block: try_wait suspends in hardware)
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred p;\n"
      "RG_WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra RG_DONE_%=;\n bra RG_WAIT_%=;\n"
      "RG_DONE_%=:\n}"
      ::"r"(bar), "r"(parity)
      : "memory");
}

__device__ __forceinline__ void rg_producer(const FpnDesc& d, const RgCfg& c, const RgMaps& maps, const RgWs& w,
                                            RgCtl* ctl, unsigned char* arena, int lane) {
  const int n_fb = w.hdr[1];
  const int fb_chunks = (c.C + 31) / 32;
  const int total = c.n_units + n_fb * fb_chunks;
  uint32_t u = 0;               // descriptors published
  int q = 0;                    // fills issued
  int jbase = 0, last_lvl = -1, done = 0;     // done: every job below this number is known to be finished
  long long t_desc = 0, t_slot = 0, t_all = clock64();
  auto publish = [&](const RgDesc& ds) {
    const int slot = u % kRgDesc;
    { RG_T0(); mbar_wait(&ctl->dempty[slot], ((u / kRgDesc) & 1u) ^ 1u); RG_T1(t_desc); }
    if (lane == 0) {
      ctl->desc[slot] = ds;
      mbar_arrive(&ctl->dfull[slot]);
    }
    ++u;
  };
  auto wait_jobs = [&](int t) {       // every job numbered below t is finished
    if (done >= t) return;
    RG_T0();
    for (;;) {
      const int p = lane < kRgWarps ? lds_acquire(&ctl->prog[lane]) : 0x7fffffff;
      done = __reduce_min_sync(0xffffffffu, p);
      if (done >= t) break;
      __nanosleep(32);
    }
    RG_T1(t_slot);
  };
  RgDesc ds;
  for (;;) {
    int unit = 0;
    if (lane == 0) unit = atomicAdd(&w.hdr[0], 1);
    unit = __shfl_sync(0xffffffffu, unit, 0);
    if (unit >= total) break;
    ds.kind = 0; ds.ncur = ds.c0 = ds.jobs = ds.tstart = ds.jbase = ds.ka = ds.kb = 0;
    ds.qoff = ds.chan_stride = ds.pitch = ds.bufoff = ds.fb_roi = ds.fb_c0 = ds.kbase = ds.mode_cls = 0;
    if (unit >= c.n_units) {
      const int f = unit - c.n_units;
      ds.kind = 1; ds.fb_roi = w.fb_list[f / fb_chunks]; ds.fb_c0 = (f % fb_chunks) * 32;
      ds.jobs = kRgWarps; ds.jbase = jbase; jbase += kRgWarps;
      publish(ds);
      continue;
    }
    int l = c.L - 1;                      // unit_base decreases with the level index
    while (l > 0 && unit >= c.lv[l - 1].unit_base) --l;
    const RgLevel& v = c.lv[l];
    const int r = unit - v.unit_base;
    const int cc = r % v.ncc, img = r / v.ncc;
    const int g0 = img * c.groups_per_img + v.group_base;
    // lane-parallel scan of the unit's groups (one round of global loads per 32 groups): jobs, first non-empty
    // group, last row any RoI needs; gneed[k] = last chunk needed by a group <= k, gcum[k] = jobs in groups <= k
    int jobs = 0, ka = 0x7fffffff, rl = -1, run_need = -1;
    for (int q0 = 0; q0 < v.nk; q0 += 32) {
      const int gq = q0 + lane;
      int ce = -1, cn = 0;
      if (gq < v.nk) {
        cn = w.cnt[g0 + gq];
        if (cn > 0) { ka = min(ka, gq); const int rm = w.rmax[g0 + gq]; rl = max(rl, rm); ce = rm / kRgCH; }
      }
      int cum = cn;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, ce, o), b = __shfl_up_sync(0xffffffffu, cum, o);
        if (lane >= o) { ce = max(ce, a); cum += b; }
      }
      ce = max(ce, run_need); cum += jobs;
      if (gq < v.nk) { ctl->gneed[gq] = ce; ctl->gcum[gq] = cum; }
      run_need = __shfl_sync(0xffffffffu, ce, 31);
      jobs = __shfl_sync(0xffffffffu, cum, 31);
    }
    __syncwarp();
    if (jobs == 0) continue;
    ka = __reduce_min_sync(0xffffffffu, ka);
    rl = __reduce_max_sync(0xffffffffu, rl);
    if (l != last_lvl) {                  // the arena changes geometry: every earlier job must be done
      wait_jobs(jbase);
      last_lvl = l;
    }
    const int c0 = cc * v.cg, ncur = min(v.cg, c.C - c0);
    const size_t plane = (size_t)v.H * v.W;
    const float* src0 = d.feat[l] + ((size_t)img * c.C + c0) * plane;
    ds.ncur = ncur; ds.c0 = c0; ds.jobs = jobs; ds.tstart = w.start[g0 + ka]; ds.jbase = jbase;
    ds.chan_stride = v.region; ds.pitch = v.pitch; ds.mode_cls = v.mode | (v.cls << 8);
    if (v.mode == 2) {                    // tall: the unit is one fill of one of two slots
      const int s = q & 1;
      ds.ka = ds.kb = 0; ds.kbase = 0; ds.qoff = q; ds.bufoff = s * v.slot_bytes;
      publish(ds);
      wait_jobs(ctl->thresh[s]);
      if (lane == 0) {
        const uint32_t bytes = (uint32_t)(ncur * plane * 4);
        u64* bar = &ctl->full[q & (kRgSlotsMax - 1)];
        mbar_arrive_expect_tx(bar, bytes);
        bulk_g2s(arena + ds.bufoff, src0, bytes, bar);
        ctl->thresh[s] = jbase + jobs;
      }
      __syncwarp();
      ++q;
      jbase += jobs;
      continue;
    }
    const int kb = min(rl / kRgCH, v.nk - 1);
    ds.ka = ka; ds.kb = kb; ds.qoff = q - ka;
    publish(ds);
    const uint32_t pitch = (uint32_t)v.pitch;
    const int ybase = (img * c.C + c0) * v.H;         // row of the tensor map: ((img * C + c) * H + y)
    int s = ka % v.NS;
    for (int k = ka; k <= kb; ++k, ++q, s = (s + 1 == v.NS) ? 0 : s + 1) {
      const int need = ctl->gneed[k];
      u64* bar = &ctl->full[q & (kRgSlotsMax - 1)];
      if (need < k) {                     // no RoI reads this chunk: an empty fill keeps the numbering in step
        if (lane == 0) mbar_arrive(bar);
        continue;
      }
      wait_jobs(ctl->thresh[s]);
      const int y0 = k * kRgCH;
      const bool wrap = s == 0 && k > 0;       // ring row NS*CH mirrors ring row 0
      if (lane == 0) {
        mbar_arrive_expect_tx(bar, (uint32_t)ncur * (kRgCH * pitch + (wrap ? (uint32_t)v.W * 4u : 0u)));
        ctl->thresh[s] = jbase + ctl->gcum[k];    // the slot is free again when the jobs of the groups <= k are done
      }
      __syncwarp();
      if (lane < ncur) {                        // lane j copies channel j: one instruction issues the whole chunk
        unsigned char* dst = arena + (size_t)lane * v.region;
        tma_2d_g2s(dst + (size_t)s * kRgCH * pitch, &maps.m[l], 0, ybase + lane * v.H + y0, bar);
        if (wrap)
          bulk_g2s(dst + (size_t)v.NS * kRgCH * pitch, src0 + (size_t)lane * plane + (size_t)y0 * v.W,
                   (uint32_t)v.W * 4u, bar);
      }
    }
    jbase += jobs;
  }
  ds.kind = 2; ds.ncur = ds.c0 = ds.jobs = ds.tstart = ds.jbase = ds.ka = ds.kb = 0;
  ds.qoff = ds.chan_stride = ds.pitch = ds.bufoff = ds.fb_roi = ds.fb_c0 = ds.kbase = ds.mode_cls = 0;
  publish(ds);
#ifdef MXD_RING_PROF
  if (lane == 0) {
    atomicAdd(reinterpret_cast<unsigned long long*>(w.hdr + 16), (unsigned long long)t_desc);
    atomicAdd(reinterpret_cast<unsigned long long*>(w.hdr + 18), (unsigned long long)t_slot);
    atomicAdd(reinterpret_cast<unsigned long long*>(w.hdr + 20), (unsigned long long)(clock64() - t_all));
  }
#else
  (void)t_desc; (void)t_slot; (void)t_all;
#endif
}

// One (RoI, unit) job of a consumer warp.  Lane = x tap (sample lane >> 1, low / high tap lane & 1).  The RoI's table
// is expanded ONCE: 14 shared-memory addresses (the low tap row of the two sample rows of every bin row, the lane's
// x offset folded in) and 28 weights that already carry the lane's x weight and 1/count.  PITCH > 0 (ring classes):
// the high tap row (+PITCH) and the channel (+jc * region) are IMMEDIATES of the LDS, so a (RoI, channel) is 28 LDS,
// 28 FFMA, a 22-instruction transposing fold and two stores, with no address arithmetic.  PITCH == 0 (tall mode):
// run-time pitch and channel stride.
template <int CG, int PITCH>
__device__ __forceinline__ void rg_job(const uint4* slot4, uint32_t px, float wx, uint32_t pitch, uint32_t chan_stride,
                                       int ncur, bool odd, bool up, bool st0, bool st1, float* o) {
  uint32_t a0[7], b0[7];
  float w00[7], w01[7], w10[7], w11[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const uint4 e = slot4[i];     // samples 2ph, 2ph+1: {row offset, l} each
    a0[i] = px + e.x; b0[i] = px + e.z;
    const float la = __uint_as_float(e.y), lb = __uint_as_float(e.w);
    w01[i] = la * wx; w00[i] = fmaf(-la, wx, wx);
    w11[i] = lb * wx; w10[i] = fmaf(-lb, wx, wx);
  }
  auto channel = [&](uint32_t joff, uint32_t hoff, float* oc) {
    float acc[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      const float v00 = lds_f32(a0[i] + joff), v01 = lds_f32(a0[i] + hoff);
      const float v10 = lds_f32(b0[i] + joff), v11 = lds_f32(b0[i] + hoff);
      acc[i] = fmaf(w11[i], v11, fmaf(w10[i], v10, fmaf(w01[i], v01, w00[i] * v00)));
    }
    // fold the 4 lanes of every bin with a transposing butterfly; lane (pw, t4) ends with bin rows t4 and t4 + 4.
    // (Keeping the rows in lane-permuted slots would make the selects disappear, but then the lanes of ONE LDS read
    // four different rows: measured 26 M extra shared-memory wavefronts on 37 M.)
    const float q0 = (odd ? acc[1] : acc[0]) + __shfl_xor_sync(0xffffffffu, odd ? acc[0] : acc[1], 1);
    const float q1 = (odd ? acc[3] : acc[2]) + __shfl_xor_sync(0xffffffffu, odd ? acc[2] : acc[3], 1);
    const float q2 = (odd ? acc[5] : acc[4]) + __shfl_xor_sync(0xffffffffu, odd ? acc[4] : acc[5], 1);
    const float q3 = acc[6] + __shfl_xor_sync(0xffffffffu, acc[6], 1);
    const float s0 = (up ? q1 : q0) + __shfl_xor_sync(0xffffffffu, up ? q0 : q1, 2);
    const float s1 = (up ? q3 : q2) + __shfl_xor_sync(0xffffffffu, up ? q2 : q3, 2);
    if (st0) oc[0] = s0;
    if (st1) oc[4 * 7] = s1;
  };
  if constexpr (PITCH == 0) {
    uint32_t joff = 0;
    for (int jc = 0; jc < ncur; ++jc) {
      channel(joff, joff + pitch, o);
      joff += chan_stride;
      o += 49;
    }
  } else {
    if (ncur == CG) {
#pragma unroll
      for (int jc = 0; jc < CG; ++jc)
        channel((uint32_t)(jc * rg_region(CG)), (uint32_t)(jc * rg_region(CG) + PITCH), o + jc * 49);
    } else {      // last channel group of a map whose C is not a multiple of cg
      for (int jc = 0; jc < ncur; ++jc)
        channel((uint32_t)(jc * rg_region(CG)), (uint32_t)(jc * rg_region(CG) + PITCH), o + jc * 49);
    }
  }
}

__global__ void __launch_bounds__(kRgThreads, 1)
roi_align_ring_fwd_kernel(const __grid_constant__ FpnDesc d, const __grid_constant__ RgCfg c,
                          const __grid_constant__ RgMaps maps, RgWs w,
                          const float* __restrict__ rois, const int* __restrict__ levels,
                          float* __restrict__ out) {
  constexpr int PW = 7, BINS = 49;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* arena = smem;
  unsigned char* tabslots = smem + kRgArena;
  RgCtl* ctl = reinterpret_cast<RgCtl*>(smem + kRgArena + kRgWarps * kRgTabSlot);
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < kRgSlotsMax; ++i) mbar_init(&ctl->full[i], 1);
    for (int i = 0; i < kRgDesc; ++i) {
      mbar_init(&ctl->dfull[i], 1);
      mbar_init(&ctl->dempty[i], kRgWarps);
    }
    fence_mbar_init();
  }
  if (tid < 32) ctl->prog[tid] = tid;       // warp w starts with job w
  if (tid < kRgSlotsMax) ctl->thresh[tid] = 0;
  __syncthreads();
  pdl_wait();                               // the planner's tables and group counts are complete
  const int lane = tid & 31, cw = tid >> 5;
  if (cw == kRgWarps) {
    rg_producer(d, c, maps, w, ctl, arena, lane);
    return;
  }
  // ================================= consumer warps ====================================
  uint2* slot = reinterpret_cast<uint2*>(tabslots + cw * kRgTabSlot);
  const uint4* slot4 = reinterpret_cast<const uint4*>(slot);
  const int xs = min(lane >> 1, 13);             // x sample of this lane
  const bool lane_on = lane < 28;
  const int t4 = lane & 3, pw = min(lane >> 2, PW - 1);
  const bool odd = lane & 1, up = lane & 2;
  const int o0 = t4 * PW + pw;                   // bin rows t4 (and t4 + 4: 28 floats further) of bin column pw
  const bool st0 = lane_on, st1 = lane_on && t4 < 3;
  const uint32_t arena_s = smem_u32(arena) + (odd ? 4u : 0u);
  const uint32_t full_s = smem_u32(&ctl->full[0]);
  uint32_t u = 0;
  int q = cw;                                    // global number of this warp's next job
  long long t_cd = 0, t_cw = 0, t_call = clock64();
  for (;;) {
    const int dslot = u % kRgDesc;
    { RG_T0(); mbar_wait(&ctl->dfull[dslot], (u / kRgDesc) & 1u); RG_T1(t_cd); }
    const int4 da = reinterpret_cast<const int4*>(&ctl->desc[dslot])[0];
    const int4 db = reinterpret_cast<const int4*>(&ctl->desc[dslot])[1];
    const int4 dc = reinterpret_cast<const int4*>(&ctl->desc[dslot])[2];
    const int4 dd = reinterpret_cast<const int4*>(&ctl->desc[dslot])[3];
    ++u;
    const int kind = da.x;
    if (kind == 2) {
#ifdef MXD_RING_PROF
      if (lane == 0) {
        atomicAdd(reinterpret_cast<unsigned long long*>(w.hdr + 22), (unsigned long long)t_cd);
        atomicAdd(reinterpret_cast<unsigned long long*>(w.hdr + 24), (unsigned long long)t_cw);
        atomicAdd(reinterpret_cast<unsigned long long*>(w.hdr + 26), (unsigned long long)(clock64() - t_call));
      }
#else
      (void)t_cd; (void)t_cw; (void)t_call;
#endif
      break;
    }
    const int jobs = da.w, jbase = db.y;
    const int jend = jbase + jobs;
    if (kind == 1) {
      if (q < jend) {       // one share of the gather per warp
        const RoiGeom g = roi_geom(d, rois, levels, dd.x, 7, 7, 2, c.finest);
        gather_roi_chunk<false>(d, g, dd.x, dd.y, min(32, c.C - dd.y), out, 7, 7, (q - jbase) * 32 + lane, kRgWarps * 32);
        q += kRgWarps;
        __syncwarp();
        if (lane == 0) sts_release(&ctl->prog[cw], q);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->dempty[dslot]);
      continue;
    }
    const int ncur = da.y, c0 = da.z;
    const int tstart = db.x - jbase;             // job q's table is tabg[tstart + q]
    const int qoff = dc.x;
    const uint32_t chan_stride = (uint32_t)dc.y, pitch = (uint32_t)dc.z;
    const uint32_t bufbase = arena_s + (uint32_t)dc.w;
    const int mode = dd.w & 0xff, cls = (dd.w >> 8) & 0xff;
    int kwait = 0;                               // chunks below kwait have been waited for by this warp
    uint2 ye = make_uint2(0u, 0u), xe = make_uint2(0u, 0u), he = make_uint2(0u, 0u);
    if (q < jend) {
      const uint2* t = w.tabg + (size_t)(tstart + q) * kRgEnt;
      if (lane < 14) ye = t[lane];
      xe = t[14 + xs];
      he = t[28];
    }
    while (q < jend) {
      const int n = (int)he.x;
      const int kf = (int)(he.y & 0xffffu), ce = (int)(he.y >> 16);
      __syncwarp();                         // the previous job's reads of the slot are done
      if (lane < 14) slot[lane] = ye;
      const uint2 xc = xe;
      __syncwarp();
      {   // next job's table: in flight during this job
        const int qn = q + kRgWarps;
        if (qn < jend) {
          const uint2* t = w.tabg + (size_t)(tstart + qn) * kRgEnt;
          if (lane < 14) ye = t[lane];
          xe = t[14 + xs];
          he = t[28];
        }
      }
      {   // the chunks this job reads and this warp has not seen yet
        RG_T0();
#ifndef MXD_RING_NOWAIT                   // experiment: pool whatever is in shared memory, never wait for the ring
        for (int k = max(kwait, kf); k <= ce; ++k) {
          const int f = qoff + k;
          mbar_wait_addr(full_s + (uint32_t)(f & (kRgSlotsMax - 1)) * 8u, (uint32_t)(f >> 4) & 1u);
        }
#endif
        kwait = ce + 1;
        RG_T1(t_cw);
      }
      float* o = out + ((size_t)n * c.C + c0) * BINS;
      const uint32_t px = bufbase + xc.x;
      const float lx = __uint_as_float(xc.y);
      const float wx = lane_on ? (odd ? lx : 1.0f - lx) * 0.25f : 0.0f;
#ifdef MXD_RING_NOCOMPUTE
      if (lx > 1e30f) o[0] = wx + __uint_as_float(px);       // experiment: stream the maps, pool nothing
      if (lx > 1e30f)
#endif
      if (mode == 1) {
        switch (cls) {
          case 0: rg_job<rg_cls_cg(0), rg_cls_pitch(0)>(slot4, px, wx, 0u, 0u, ncur, odd, up, st0, st1, o + o0); break;
          case 1: rg_job<rg_cls_cg(1), rg_cls_pitch(1)>(slot4, px, wx, 0u, 0u, ncur, odd, up, st0, st1, o + o0); break;
          case 2: rg_job<rg_cls_cg(2), rg_cls_pitch(2)>(slot4, px, wx, 0u, 0u, ncur, odd, up, st0, st1, o + o0); break;
          default: rg_job<rg_cls_cg(3), rg_cls_pitch(3)>(slot4, px, wx, 0u, 0u, ncur, odd, up, st0, st1, o + o0); break;
        }
      } else {
        rg_job<1, 0>(slot4, px, wx, pitch, chan_stride, ncur, odd, up, st0, st1, o + o0);
      }
      q += kRgWarps;
      __syncwarp();                         // every lane has its values: the chunks may be overwritten
      if (lane == 0) sts_release(&ctl->prog[cw], q);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&ctl->dempty[dslot]);
  }
}

size_t ring_workspace_bytes(int R, int N, int L, const int* Hs, const int* Ws, int C, int PH, int PW, int sr) {
  FpnDesc d = {};
  d.num_levels = L; d.N = N; d.C = C;
  for (int l = 0; l < L; ++l) { d.H[l] = Hs[l]; d.W[l] = Ws[l]; }      // null pointers count as aligned
  RgCfg c;
  if (!rg_make_cfg(d, PH, PW, sr, 56.0f, &c) || c.NB > kRgMaxGroups) return 0;
  return rg_carve(nullptr, R, c.NB).bytes;
}

#ifdef MXD_RING_PROF
extern "C" int mxd_ring_prof(const void* ws, unsigned long long* out6) {
  return (int)cudaMemcpy(out6, (const char*)ws + 64, 48, cudaMemcpyDeviceToHost);
}
#endif

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point (libcuda is not linked).
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

int ring_forward(const FpnDesc& d, const float* rois, const int* levels, float* out, int R, int PH, int PW, int sr,
                 float finest, void* ws, size_t ws_bytes, cudaStream_t st, int* handled) {
  *handled = 0;
  RgCfg c;
  if (R == 0 || d.C == 0) return MXD_OK;
  if (getenv("MXD_NO_RING") != nullptr) return MXD_OK;      // A/B and test switch: the band kernels take the call
  EncodeTiledFn enc = encode_tiled_fn();
  if (!rg_make_cfg(d, PH, PW, sr, finest, &c, enc != nullptr) || c.NB > kRgMaxGroups) return MXD_OK;
  int sms = 0, dev = 0;
  MXD_CUDA_OK(cudaGetDevice(&dev));
  MXD_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  // a unit is a whole (image, channel group) plane pass: with fewer than two units per SM (one small image) the band
  // kernels, whose items are row bands, balance better
  if (c.n_units < 2 * sms && getenv("MXD_RING_FORCE") == nullptr) return MXD_OK;      // (tests force the ring path)
  RgMaps maps;
  memset(&maps, 0, sizeof(maps));
  for (int l = 0; l < c.L; ++l) {
    const RgLevel& v = c.lv[l];
    if (v.mode != 1) continue;
    // the map as (W/2) x (N*C*H) 8-byte elements: a box of pitch/8 elements x CH rows; columns past W read as zero
    const cuuint64_t gdim[2] = {(cuuint64_t)v.W / 2, (cuuint64_t)d.N * d.C * v.H};
    const cuuint64_t gstr[1] = {(cuuint64_t)v.W * 4};
    const cuuint32_t box[2] = {(cuuint32_t)v.pitch / 8, (cuuint32_t)kRgCH};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&maps.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, d.feat[l], gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return MXD_OK;       // not encodable (e.g. > 2^32 rows): the band kernels take the call
  }
  RgWs w = rg_carve(ws, R, c.NB);
  MXD_REQUIRE(ws_bytes >= w.bytes, MXD_EWORKSPACE, "roi_align workspace %zu < %zu bytes", ws_bytes, w.bytes);
  MXD_REQUIRE(((uintptr_t)ws & 255) == 0, MXD_EINVAL, "workspace must be 256-byte aligned");
  const size_t zbytes = (size_t)((char*)w.start - (char*)w.hdr);      // hdr, cnt, rmax
  MXD_CUDA_OK(cudaMemsetAsync(w.hdr, 0, zbytes, st));
  MXD_CUDA_OK(launch_pdl(rg_plan_rois_kernel, dim3((R * 32 + 255) / 256), dim3(256), 0, st, d, c, w, rois, levels, R));
  MXD_POST_LAUNCH("roi_align_rg_plan_rois");
  MXD_CUDA_OK(launch_pdl(rg_plan_pack_kernel, dim3((R + 7) / 8), dim3(256), 0, st, c, w, R));
  MXD_POST_LAUNCH("roi_align_rg_plan_pack");
  static unsigned long long seen = 0;
  DeviceOnce once_seen(&seen);
  if (once_seen.first())
    MXD_CUDA_OK(cudaFuncSetAttribute(roi_align_ring_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRgSmem));
  MXD_CUDA_OK(launch_pdl(roi_align_ring_fwd_kernel, dim3(sms), dim3(kRgThreads), (size_t)kRgSmem, st, d, c, maps, w, rois,
                         levels, out));
  MXD_POST_LAUNCH("roi_align_ring_fwd");
  *handled = 1;
  return MXD_OK;
}

}  // namespace mxd
