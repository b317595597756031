// Shared device code of the RoIAlign kernels (Spec A / Spec G).
#pragma once
#include "common.cuh"

namespace mxd {

struct FpnDesc {
  float* feat[MXD_MAX_LEVELS];
  int H[MXD_MAX_LEVELS];
  int W[MXD_MAX_LEVELS];
  float scale[MXD_MAX_LEVELS];
  int num_levels;
  int N, C;
};

struct AxisTap {
  int lo, hi;      // element offsets along the axis (already multiplied by the pitch)
  float l, h;      // weights of the hi / lo taps
  int valid;
};

// One sample coordinate of Spec A along one axis.  `pitch` = 1 for x, W for y.
__device__ __forceinline__ AxisTap axis_tap(float start, float bin, int grid, int p, int i, int size, int pitch) {
  AxisTap t;
  // c = (start + p*bin) + ((i+.5f)*bin)/grid   -- order of Spec A
  float c = __fadd_rn(__fadd_rn(start, __fmul_rn((float)p, bin)),
                      __fdiv_rn(__fmul_rn(__fadd_rn((float)i, 0.5f), bin), (float)grid));
  t.valid = !(c < -1.0f || c > (float)size);
  if (c <= 0.0f) c = 0.0f;
  int lo = t.valid ? (int)c : 0;
  int hi;
  if (lo >= size - 1) {
    hi = lo = size - 1;
    c = (float)lo;
  } else {
    hi = lo + 1;
  }
  t.l = __fsub_rn(c, (float)lo);
  t.h = __fsub_rn(1.0f, t.l);
  t.lo = lo * pitch;
  t.hi = hi * pitch;
  return t;
}

struct RoiGeom {
  int b, lvl, H, W, gh, gw;
  float rsw, rsh, bh, bw;
  float* plane0;  // feat[lvl] + b*C*H*W
  bool ok;
};

__device__ __forceinline__ RoiGeom roi_geom(const FpnDesc& d, const float* __restrict__ rois,
                                            const int* __restrict__ levels, int n, int PH, int PW, int sr,
                                            float finest) {
  RoiGeom g;
  const float* r = rois + (size_t)n * 5;
  float rb = r[0], x1 = r[1], y1 = r[2], x2 = r[3], y2 = r[4];
  g.b = (int)rb;
  g.lvl = 0;
  if (d.num_levels > 1) g.lvl = levels ? levels[n] : roi_level(x1, y1, x2, y2, d.num_levels, finest);
  g.ok = g.b >= 0 && g.b < d.N && g.lvl >= 0 && g.lvl < d.num_levels;
  int lv = g.ok ? g.lvl : 0;
  g.H = d.H[lv];
  g.W = d.W[lv];
  float sc = d.scale[lv];
  g.rsw = __fmul_rn(x1, sc);
  g.rsh = __fmul_rn(y1, sc);
  float rew = __fmul_rn(x2, sc), reh = __fmul_rn(y2, sc);
  float rw = fmaxf(__fsub_rn(rew, g.rsw), 1.0f), rh = fmaxf(__fsub_rn(reh, g.rsh), 1.0f);
  g.bh = __fdiv_rn(rh, (float)PH);
  g.bw = __fdiv_rn(rw, (float)PW);
  g.gh = sr > 0 ? sr : (int)ceilf(g.bh);
  g.gw = sr > 0 ? sr : (int)ceilf(g.bw);
  g.plane0 = d.feat[lv] + (size_t)(g.ok ? g.b : 0) * d.C * g.H * g.W;
  return g;
}

// Generic per-RoI gather: outputs (c in [c0,c0+cc), bin) of RoI n, taps read straight from global
// memory (forward) or scattered with RED.ADD (backward).  Used by the v1 kernels and as the in-kernel
// fallback of the plane-resident kernels for RoIs that do not fit a shared-memory window.
template <bool BWD>
__device__ __forceinline__ void gather_roi_chunk(const FpnDesc& d, const RoiGeom& g, int n, int c0, int cc,
                                                 float* __restrict__ io, int PH, int PW, int tid, int nthreads) {
  const int bins = PH * PW;
  float* io_base = io + ((size_t)n * d.C + c0) * bins;
  if (!g.ok) {  // Spec A: batch index (or level) out of range -> zeros / no gradient
    if (!BWD)
      for (int o = tid; o < cc * bins; o += nthreads) io_base[o] = 0.0f;
    return;
  }
  const float count = (float)(g.gh * g.gw);
  const float inv_count = 1.0f / count;
  const size_t plane_sz = (size_t)g.H * g.W;
  for (int o = tid; o < cc * bins; o += nthreads) {
    const int c = o / bins;
    const int bin = o - c * bins;
    const int ph = bin / PW, pw = bin - ph * PW;
    float* plane = g.plane0 + (size_t)(c0 + c) * plane_sz;
    float acc = 0.0f;
    float gscaled = 0.0f;
    if (BWD) gscaled = io_base[o] * inv_count;
    for (int iy = 0; iy < g.gh; ++iy) {
      const AxisTap y = axis_tap(g.rsh, g.bh, g.gh, ph, iy, g.H, g.W);
      if (!y.valid) continue;
      for (int ix = 0; ix < g.gw; ++ix) {
        const AxisTap x = axis_tap(g.rsw, g.bw, g.gw, pw, ix, g.W, 1);
        if (!x.valid) continue;
        const float w1 = y.h * x.h, w2 = y.h * x.l, w3 = y.l * x.h, w4 = y.l * x.l;
        if (!BWD) {
          const float v1 = __ldg(plane + y.lo + x.lo), v2 = __ldg(plane + y.lo + x.hi);
          const float v3 = __ldg(plane + y.hi + x.lo), v4 = __ldg(plane + y.hi + x.hi);
          acc += ((w1 * v1 + w2 * v2) + w3 * v3) + w4 * v4;
        } else {
          atomicAdd(plane + y.lo + x.lo, gscaled * w1);
          atomicAdd(plane + y.lo + x.hi, gscaled * w2);
          atomicAdd(plane + y.hi + x.lo, gscaled * w3);
          atomicAdd(plane + y.hi + x.hi, gscaled * w4);
        }
      }
    }
    if (!BWD) io_base[o] = acc / count;
  }
}

// ---- plane-resident path (roi_align_plane.cu) -------------------------------------
// Returns MXD_OK and sets *handled=1 when the plane kernels ran; *handled=0 means the
// configuration is outside their range and the caller should use the gather kernels.
size_t plane_workspace_bytes(int R, int N, int L, const int* Hs, const int* Ws, int C, int PH, int PW, int sr);
int plane_forward(const FpnDesc& d, const float* rois, const int* levels, float* out, int R, int PH, int PW,
                  int sr, float finest, void* ws, size_t ws_bytes, cudaStream_t st, int* handled);
int plane_backward(const FpnDesc& d, const float* rois, const int* levels, const float* gout, int R, int PH,
                   int PW, int sr, float finest, int accumulate, void* ws, size_t ws_bytes, cudaStream_t st,
                   int* handled);

// ---- row-ring forward (roi_align_ring.cu): same convention -------------------------
size_t ring_workspace_bytes(int R, int N, int L, const int* Hs, const int* Ws, int C, int PH, int PW, int sr);
int ring_forward(const FpnDesc& d, const float* rois, const int* levels, float* out, int R, int PH, int PW, int sr,
                 float finest, void* ws, size_t ws_bytes, cudaStream_t st, int* handled);

// ---- tile-resident backward (roi_align_tile_bwd.cu): same convention --------------
size_t tile_bwd_workspace_bytes(int R, int N, int L, const int* Hs, const int* Ws, int C, int PH, int PW, int sr);
int tile_backward(const FpnDesc& d, const float* rois, const int* levels, const float* gout, int R, int PH,
                  int PW, int sr, float finest, int accumulate, void* ws, size_t ws_bytes, cudaStream_t st,
                  int* handled);

}  // namespace mxd
