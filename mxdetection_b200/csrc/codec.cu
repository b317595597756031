// Box delta encode / decode / clip (Spec F, rows F1/F2).
// Module role: mxdetection/core/bbox (/root/reference/README.md:17);
// bbox2delta / delta2bbox of mmdet 0.5.
#include "internal.h"

namespace mxd {

struct F4 { float v[4]; };

__global__ void bbox2delta_kernel(const float4* __restrict__ p, const float4* __restrict__ g, int m, F4 means,
                                  F4 stds, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  out[i] = encode_box(p[i], g[i], means.v, stds.v);
}

__global__ void delta2bbox_kernel(const float4* __restrict__ r, const float4* __restrict__ d, int m, F4 means,
                                  F4 stds, float max_ratio, float hmax, float wmax, int clip,
                                  float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  out[i] = decode_box(r[i], d[i], means.v, stds.v, max_ratio, hmax, wmax, clip != 0);
}

static int check_m4(const DLTensor* t, const char* name, long long m, int* dev) {
  int rc;
  if ((rc = check_tensor(t, name, F32, 2, 2, dev))) return rc;
  MXD_REQUIRE(t->shape[1] == 4 && (m < 0 || t->shape[0] == m), MXD_EINVAL, "%s must be (M,4)", name);
  MXD_REQUIRE(((uintptr_t)dptr<float>(t) & 15) == 0, MXD_EINVAL, "%s must be 16-byte aligned", name);
  return MXD_OK;
}

}  // namespace mxd

using namespace mxd;

extern "C" {

int mxd_bbox2delta(const DLTensor* proposals, const DLTensor* gts, DLTensor* deltas, const float* means,
                   const float* stds, void* stream) {
  int dev = -1, rc;
  if ((rc = check_m4(proposals, "proposals", -1, &dev))) return rc;
  const long long m = proposals->shape[0];
  if ((rc = check_m4(gts, "gts", m, &dev))) return rc;
  if ((rc = check_m4(deltas, "deltas", m, &dev))) return rc;
  MXD_REQUIRE(means && stds && m < (1ll << 31), MXD_EINVAL, "means/stds must be float[4]");
  if (m == 0) return MXD_OK;
  F4 mu, sd;
  for (int j = 0; j < 4; ++j) { mu.v[j] = means[j]; sd.v[j] = stds[j]; }
  bbox2delta_kernel<<<((int)m + 255) / 256, 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(dptr<float>(proposals)), reinterpret_cast<const float4*>(dptr<float>(gts)),
      (int)m, mu, sd, reinterpret_cast<float4*>(dptr<float>(deltas)));
  MXD_POST_LAUNCH("bbox2delta");
  return MXD_OK;
}

int mxd_delta2bbox(const DLTensor* rois, const DLTensor* deltas, DLTensor* boxes, const float* means,
                   const float* stds, int max_h, int max_w, double wh_ratio_clip, void* stream) {
  int dev = -1, rc;
  if ((rc = check_m4(rois, "rois", -1, &dev))) return rc;
  const long long m = rois->shape[0];
  if ((rc = check_m4(deltas, "deltas", m, &dev))) return rc;
  if ((rc = check_m4(boxes, "boxes", m, &dev))) return rc;
  MXD_REQUIRE(means && stds && m < (1ll << 31), MXD_EINVAL, "means/stds must be float[4]");
  MXD_REQUIRE(wh_ratio_clip > 0, MXD_EINVAL, "wh_ratio_clip must be > 0");
  if (m == 0) return MXD_OK;
  F4 mu, sd;
  for (int j = 0; j < 4; ++j) { mu.v[j] = means[j]; sd.v[j] = stds[j]; }
  const float max_ratio = (float)fabs(log(wh_ratio_clip));
  const int clip = (max_h > 0 && max_w > 0) ? 1 : 0;
  delta2bbox_kernel<<<((int)m + 255) / 256, 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(dptr<float>(rois)), reinterpret_cast<const float4*>(dptr<float>(deltas)),
      (int)m, mu, sd, max_ratio, (float)(max_h - 1), (float)(max_w - 1), clip,
      reinterpret_cast<float4*>(dptr<float>(boxes)));
  MXD_POST_LAUNCH("delta2bbox");
  return MXD_OK;
}

}  // extern "C"
