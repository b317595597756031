// RoIAlign forward / backward, single map and FPN multi-level (Spec A, G).
//
// Contract: mx.nd.contrib.ROIAlign / _backward_ROIAlign of mxnet 1.3.0
// (module mxdetection/ops, /root/reference/README.md:24) and
// SingleLevelRoI.forward (mxdetection/models/roi_extractors,
// /root/reference/README.md:32).  NCHW fp32.
//
// v1 "gather" kernels: one CTA per (RoI, channel chunk).  The per-axis tap
// tables (index, weights, validity) are computed once per CTA in shared
// memory - they are shared by every channel - then each thread produces
// outputs (c, ph, pw) with coalesced stores and read-only-path tap loads.
#include <algorithm>
#include "roi_align.cuh"

namespace mxd {

constexpr int kTab = 64;       // max samples per axis held in the smem tables
constexpr int kThreads = 256;

template <bool BWD, bool TAB>
__global__ void __launch_bounds__(kThreads)
roi_align_gather_kernel(FpnDesc d, const float* __restrict__ rois, const int* __restrict__ levels,
                        float* __restrict__ io, int PH, int PW, int sr, float finest, int c_chunk) {
  const int n = blockIdx.x;
  const int c0 = blockIdx.y * c_chunk;
  const int cc = min(c_chunk, d.C - c0);
  const RoiGeom g = roi_geom(d, rois, levels, n, PH, PW, sr, finest);
  const int bins = PH * PW;
  float* io_base = io + ((size_t)n * d.C + c0) * bins;

  if (!g.ok) {  // Spec A: batch index (or level) out of range -> zeros / no gradient
    if (!BWD)
      for (int o = threadIdx.x; o < cc * bins; o += kThreads) io_base[o] = 0.0f;
    return;
  }

  __shared__ int s_ylo[TAB ? kTab : 1], s_yhi[TAB ? kTab : 1], s_xlo[TAB ? kTab : 1], s_xhi[TAB ? kTab : 1];
  __shared__ float s_ly[TAB ? kTab : 1], s_hy[TAB ? kTab : 1], s_lx[TAB ? kTab : 1], s_hx[TAB ? kTab : 1];
  if (TAB) {
    for (int t = threadIdx.x; t < PH * g.gh; t += kThreads) {
      AxisTap a = axis_tap(g.rsh, g.bh, g.gh, t / g.gh, t % g.gh, g.H, g.W);
      s_ylo[t] = a.valid ? a.lo : -1;
      s_yhi[t] = a.hi; s_ly[t] = a.l; s_hy[t] = a.h;
    }
    for (int t = threadIdx.x; t < PW * g.gw; t += kThreads) {
      AxisTap a = axis_tap(g.rsw, g.bw, g.gw, t / g.gw, t % g.gw, g.W, 1);
      s_xlo[t] = a.valid ? a.lo : -1;
      s_xhi[t] = a.hi; s_lx[t] = a.l; s_hx[t] = a.h;
    }
    __syncthreads();
  }

  const float count = (float)(g.gh * g.gw);
  const float inv_count = 1.0f / count;
  const size_t plane_sz = (size_t)g.H * g.W;

  for (int o = threadIdx.x; o < cc * bins; o += kThreads) {
    const int c = o / bins;
    const int bin = o - c * bins;
    const int ph = bin / PW, pw = bin - ph * PW;
    float* plane = g.plane0 + (size_t)(c0 + c) * plane_sz;
    float acc = 0.0f;
    float gscaled = 0.0f;
    if (BWD) gscaled = io_base[o] * inv_count;
    for (int iy = 0; iy < g.gh; ++iy) {
      int ylo, yhi; float ly, hy;
      if (TAB) {
        const int t = ph * g.gh + iy;
        ylo = s_ylo[t]; yhi = s_yhi[t]; ly = s_ly[t]; hy = s_hy[t];
      } else {
        AxisTap a = axis_tap(g.rsh, g.bh, g.gh, ph, iy, g.H, g.W);
        ylo = a.valid ? a.lo : -1; yhi = a.hi; ly = a.l; hy = a.h;
      }
      if (ylo < 0) continue;
      for (int ix = 0; ix < g.gw; ++ix) {
        int xlo, xhi; float lx, hx;
        if (TAB) {
          const int t = pw * g.gw + ix;
          xlo = s_xlo[t]; xhi = s_xhi[t]; lx = s_lx[t]; hx = s_hx[t];
        } else {
          AxisTap a = axis_tap(g.rsw, g.bw, g.gw, pw, ix, g.W, 1);
          xlo = a.valid ? a.lo : -1; xhi = a.hi; lx = a.l; hx = a.h;
        }
        if (xlo < 0) continue;
        const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
        if (!BWD) {
          const float v1 = __ldg(plane + ylo + xlo), v2 = __ldg(plane + ylo + xhi);
          const float v3 = __ldg(plane + yhi + xlo), v4 = __ldg(plane + yhi + xhi);
          acc += ((w1 * v1 + w2 * v2) + w3 * v3) + w4 * v4;
        } else {
          atomicAdd(plane + ylo + xlo, gscaled * w1);
          atomicAdd(plane + ylo + xhi, gscaled * w2);
          atomicAdd(plane + yhi + xlo, gscaled * w3);
          atomicAdd(plane + yhi + xhi, gscaled * w4);
        }
      }
    }
    if (!BWD) io_base[o] = acc / count;
  }
}

template <bool BWD>
static int launch_gather(const FpnDesc& d, const float* rois, const int* levels, float* io, int R,
                         int PH, int PW, int sr, float finest, cudaStream_t st) {
  if (R == 0 || d.C == 0) return MXD_OK;
  int c_chunk = 32;
  while ((d.C + c_chunk - 1) / c_chunk > 65535) c_chunk *= 2;
  dim3 grid(R, (d.C + c_chunk - 1) / c_chunk);
  const bool tab = sr > 0 && PH * sr <= kTab && PW * sr <= kTab;
  if (tab)
    roi_align_gather_kernel<BWD, true><<<grid, kThreads, 0, st>>>(d, rois, levels, io, PH, PW, sr, finest, c_chunk);
  else
    roi_align_gather_kernel<BWD, false><<<grid, kThreads, 0, st>>>(d, rois, levels, io, PH, PW, sr, finest, c_chunk);
  MXD_POST_LAUNCH(BWD ? "roi_align_backward_gather" : "roi_align_forward_gather");
  return MXD_OK;
}

__global__ void map_levels_kernel(const float* __restrict__ rois, int cols, int R, int L, float finest,
                                  int* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  const float* r = rois + (size_t)i * cols + (cols == 5 ? 1 : 0);
  out[i] = roi_level(r[0], r[1], r[2], r[3], L, finest);
}

static int build_desc(const DLTensor* const* feats, int L, const float* scales, FpnDesc* d, int* dev,
                      const char* what) {
  MXD_REQUIRE(L >= 1 && L <= MXD_MAX_LEVELS, MXD_EINVAL, "num_levels %d not in [1,%d]", L, MXD_MAX_LEVELS);
  MXD_REQUIRE(feats && scales, MXD_EINVAL, "%s: null level table", what);
  d->num_levels = L;
  for (int l = 0; l < L; ++l) {
    int rc = check_tensor(feats[l], what, F32, 4, 4, dev);
    if (rc) return rc;
    if (l == 0) { d->N = (int)feats[l]->shape[0]; d->C = (int)feats[l]->shape[1]; }
    MXD_REQUIRE(feats[l]->shape[0] == d->N && feats[l]->shape[1] == d->C, MXD_EINVAL,
                "%s: level %d has (N,C)=(%lld,%lld), level 0 has (%d,%d)", what, l,
                (long long)feats[l]->shape[0], (long long)feats[l]->shape[1], d->N, d->C);
    MXD_REQUIRE(feats[l]->shape[2] >= 1 && feats[l]->shape[3] >= 1 &&
                feats[l]->shape[2] * feats[l]->shape[3] < (1ll << 31), MXD_EINVAL, "%s: bad H,W", what);
    d->feat[l] = dptr<float>(feats[l]);
    MXD_REQUIRE(((uintptr_t)d->feat[l] & 3) == 0, MXD_EINVAL, "%s: misaligned data", what);
    d->H[l] = (int)feats[l]->shape[2];
    d->W[l] = (int)feats[l]->shape[3];
    d->scale[l] = scales[l];
  }
  return MXD_OK;
}

static int check_common(const DLTensor* rois, const DLTensor* levels, const DLTensor* pooled, int C,
                        int PH, int PW, int* dev, int* R) {
  int rc;
  MXD_REQUIRE(PH >= 1 && PW >= 1, MXD_EINVAL, "pooled_size must be >= 1");
  if ((rc = check_tensor(rois, "rois", F32, 2, 2, dev))) return rc;
  MXD_REQUIRE(rois->shape[1] == 5, MXD_EINVAL, "rois must be (R,5) [batch,x1,y1,x2,y2]");
  *R = (int)rois->shape[0];
  if (levels) {
    if ((rc = check_tensor(levels, "levels", I32, 1, 1, dev))) return rc;
    MXD_REQUIRE(levels->shape[0] == *R, MXD_EINVAL, "levels must be (R)");
  }
  if ((rc = check_tensor(pooled, "pooled", F32, 4, 4, dev))) return rc;
  MXD_REQUIRE(pooled->shape[0] == *R && pooled->shape[1] == C && pooled->shape[2] == PH &&
              pooled->shape[3] == PW, MXD_EINVAL, "pooled tensor must be (R=%d,C=%d,%d,%d)", *R, C, PH, PW);
  return MXD_OK;
}

}  // namespace mxd

using namespace mxd;

extern "C" {

size_t mxd_roi_align_workspace_bytes(int num_rois, int batch, int channels, int num_levels, const int* feat_h,
                                     const int* feat_w, int pooled_h, int pooled_w, int sample_ratio) {
  if (num_levels < 1 || num_levels > MXD_MAX_LEVELS || !feat_h || !feat_w) return 0;
  const size_t a = plane_workspace_bytes(num_rois, batch, num_levels, feat_h, feat_w, channels, pooled_h, pooled_w, sample_ratio);
  const size_t b = tile_bwd_workspace_bytes(num_rois, batch, num_levels, feat_h, feat_w, channels, pooled_h, pooled_w, sample_ratio);
  const size_t r = ring_workspace_bytes(num_rois, batch, num_levels, feat_h, feat_w, channels, pooled_h, pooled_w, sample_ratio);
  return std::max(a, std::max(b, r));
}

int mxd_roi_align_fpn_forward(const DLTensor* const* feats, int num_levels, const float* spatial_scales,
                              const DLTensor* rois, const DLTensor* levels, DLTensor* out,
                              int pooled_h, int pooled_w, int sample_ratio, float finest_scale,
                              void* workspace, size_t workspace_bytes, void* stream) {
  FpnDesc d; int dev = -1, R = 0, rc;
  if ((rc = build_desc(feats, num_levels, spatial_scales, &d, &dev, "feats"))) return rc;
  if ((rc = check_common(rois, levels, out, d.C, pooled_h, pooled_w, &dev, &R))) return rc;
  const int* lv = levels ? dptr<int>(levels) : nullptr;
  if (workspace) {
    int handled = 0;
    if ((rc = ring_forward(d, dptr<float>(rois), lv, dptr<float>(out), R, pooled_h, pooled_w, sample_ratio,
                           finest_scale, workspace, workspace_bytes, as_stream(stream), &handled))) return rc;
    if (handled) return MXD_OK;
    if ((rc = plane_forward(d, dptr<float>(rois), lv, dptr<float>(out), R, pooled_h, pooled_w, sample_ratio,
                            finest_scale, workspace, workspace_bytes, as_stream(stream), &handled))) return rc;
    if (handled) return MXD_OK;
  }
  return launch_gather<false>(d, dptr<float>(rois), lv, dptr<float>(out), R, pooled_h, pooled_w, sample_ratio,
                              finest_scale, as_stream(stream));
}

int mxd_roi_align_fpn_backward(const DLTensor* grad_out, const DLTensor* rois, const DLTensor* levels,
                               DLTensor* const* grad_feats, int num_levels, const float* spatial_scales,
                               int pooled_h, int pooled_w, int sample_ratio, float finest_scale,
                               int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  FpnDesc d; int dev = -1, R = 0, rc;
  if ((rc = build_desc(grad_feats, num_levels, spatial_scales, &d, &dev, "grad_feats"))) return rc;
  if ((rc = check_common(rois, levels, grad_out, d.C, pooled_h, pooled_w, &dev, &R))) return rc;
  cudaStream_t st = as_stream(stream);
  const int* lv = levels ? dptr<int>(levels) : nullptr;
  if (workspace) {
    int handled = 0;
    if ((rc = tile_backward(d, dptr<float>(rois), lv, dptr<float>(grad_out), R, pooled_h, pooled_w, sample_ratio,
                            finest_scale, accumulate, workspace, workspace_bytes, st, &handled))) return rc;
    if (handled) return MXD_OK;
    if ((rc = plane_backward(d, dptr<float>(rois), lv, dptr<float>(grad_out), R, pooled_h, pooled_w, sample_ratio,
                             finest_scale, accumulate, workspace, workspace_bytes, st, &handled))) return rc;
    if (handled) return MXD_OK;
  }
  if (!accumulate)
    for (int l = 0; l < num_levels; ++l)
      MXD_CUDA_OK(cudaMemsetAsync(d.feat[l], 0, sizeof(float) * (size_t)numel(grad_feats[l]), st));
  return launch_gather<true>(d, dptr<float>(rois), lv, dptr<float>(grad_out), R, pooled_h, pooled_w, sample_ratio,
                             finest_scale, st);
}

int mxd_roi_align_forward(const DLTensor* data, const DLTensor* rois, DLTensor* out, int pooled_h,
                          int pooled_w, float spatial_scale, int sample_ratio, void* workspace,
                          size_t workspace_bytes, void* stream) {
  return mxd_roi_align_fpn_forward(&data, 1, &spatial_scale, rois, nullptr, out, pooled_h, pooled_w,
                                   sample_ratio, 56.0f, workspace, workspace_bytes, stream);
}

int mxd_roi_align_backward(const DLTensor* grad_out, const DLTensor* rois, DLTensor* grad_data,
                           int pooled_h, int pooled_w, float spatial_scale, int sample_ratio,
                           int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  return mxd_roi_align_fpn_backward(grad_out, rois, nullptr, &grad_data, 1, &spatial_scale, pooled_h,
                                    pooled_w, sample_ratio, 56.0f, accumulate, workspace, workspace_bytes, stream);
}

int mxd_map_roi_levels(const DLTensor* rois, DLTensor* levels, int num_levels, float finest_scale,
                       void* stream) {
  int dev = -1, rc;
  if ((rc = check_tensor(rois, "rois", F32, 2, 2, &dev))) return rc;
  MXD_REQUIRE(rois->shape[1] == 5 || rois->shape[1] == 4, MXD_EINVAL, "rois must be (R,5) or (R,4)");
  if ((rc = check_tensor(levels, "levels", I32, 1, 1, &dev))) return rc;
  MXD_REQUIRE(levels->shape[0] == rois->shape[0], MXD_EINVAL, "levels must be (R)");
  MXD_REQUIRE(num_levels >= 1 && finest_scale > 0, MXD_EINVAL, "bad num_levels / finest_scale");
  int R = (int)rois->shape[0];
  if (R == 0) return MXD_OK;
  map_levels_kernel<<<(R + 255) / 256, 256, 0, as_stream(stream)>>>(dptr<float>(rois), (int)rois->shape[1],
                                                                     R, num_levels, finest_scale, dptr<int>(levels));
  MXD_POST_LAUNCH("map_levels");
  return MXD_OK;
}

}  // extern "C"
