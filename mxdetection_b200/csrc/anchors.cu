// Anchor grid + validity flags (Spec C, rows C1/C2).
// Module role: mxdetection/core/anchor (/root/reference/README.md:16);
// AnchorGenerator.grid_anchors / valid_flags and anchor_inside_flags of mmdet 0.5.
#include "common.cuh"

namespace mxd {

struct BaseAnchors { float v[MXD_MAX_BASE_ANCHORS][4]; };

__global__ void grid_anchors_kernel(BaseAnchors base, int A, int H, int W, float stride, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W * A) return;
  const int a = i % A, cell = i / A;
  const int x = cell % W, y = cell / W;
  const float sx = __fmul_rn((float)x, stride), sy = __fmul_rn((float)y, stride);
  out[i] = make_float4(__fadd_rn(base.v[a][0], sx), __fadd_rn(base.v[a][1], sy),
                       __fadd_rn(base.v[a][2], sx), __fadd_rn(base.v[a][3], sy));
}

__global__ void valid_flags_kernel(int A, int H, int W, int vh, int vw, uint8_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W * A) return;
  const int cell = i / A;
  const int x = cell % W, y = cell / W;
  out[i] = (x < vw && y < vh) ? 1 : 0;
}

__global__ void inside_flags_kernel(const float4* __restrict__ anchors, const uint8_t* __restrict__ valid, int n,
                                    float img_h, float img_w, float ab, uint8_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool v = valid ? valid[i] != 0 : true;
  if (ab >= 0.0f) {
    const float4 a = anchors[i];
    v = v && a.x >= -ab && a.y >= -ab && a.z < __fadd_rn(img_w, ab) && a.w < __fadd_rn(img_h, ab);
  }
  out[i] = v ? 1 : 0;
}

}  // namespace mxd

using namespace mxd;

extern "C" {

int mxd_grid_anchors(const float* base_anchors, int num_base, int feat_h, int feat_w, float stride,
                     DLTensor* out, void* stream) {
  int dev = -1, rc;
  MXD_REQUIRE(base_anchors && num_base >= 1 && num_base <= MXD_MAX_BASE_ANCHORS, MXD_EINVAL,
              "num_base %d not in [1,%d]", num_base, MXD_MAX_BASE_ANCHORS);
  MXD_REQUIRE(feat_h >= 0 && feat_w >= 0 && (long long)feat_h * feat_w * num_base < (1ll << 31), MXD_EINVAL, "bad grid");
  if ((rc = check_tensor(out, "out", F32, 2, 2, &dev))) return rc;
  const int n = feat_h * feat_w * num_base;
  MXD_REQUIRE(out->shape[0] == n && out->shape[1] == 4, MXD_EINVAL, "out must be (%d,4)", n);
  MXD_REQUIRE(((uintptr_t)dptr<float>(out) & 15) == 0, MXD_EINVAL, "out must be 16-byte aligned");
  if (n == 0) return MXD_OK;
  BaseAnchors b;
  for (int a = 0; a < num_base; ++a)
    for (int j = 0; j < 4; ++j) b.v[a][j] = base_anchors[a * 4 + j];
  grid_anchors_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(b, num_base, feat_h, feat_w, stride,
                                                                       reinterpret_cast<float4*>(dptr<float>(out)));
  MXD_POST_LAUNCH("grid_anchors");
  return MXD_OK;
}

int mxd_valid_flags(int feat_h, int feat_w, int valid_h, int valid_w, int num_base, DLTensor* flags, void* stream) {
  int dev = -1, rc;
  if ((rc = check_tensor(flags, "flags", U8, 1, 1, &dev))) return rc;
  MXD_REQUIRE(feat_h >= 0 && feat_w >= 0 && num_base >= 1 && (long long)feat_h * feat_w * num_base < (1ll << 31),
              MXD_EINVAL, "bad grid");
  const int n = feat_h * feat_w * num_base;
  MXD_REQUIRE(flags->shape[0] == n, MXD_EINVAL, "flags must be (%d)", n);
  if (n == 0) return MXD_OK;
  valid_flags_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(num_base, feat_h, feat_w, valid_h, valid_w,
                                                                      dptr<uint8_t>(flags));
  MXD_POST_LAUNCH("valid_flags");
  return MXD_OK;
}

int mxd_inside_flags(const DLTensor* anchors, const DLTensor* valid, int img_h, int img_w, float allowed_border,
                     DLTensor* out, void* stream) {
  int dev = -1, rc;
  if ((rc = check_tensor(anchors, "anchors", F32, 2, 2, &dev))) return rc;
  MXD_REQUIRE(anchors->shape[1] == 4, MXD_EINVAL, "anchors must be (N,4)");
  const long long n = anchors->shape[0];
  MXD_REQUIRE(n < (1ll << 31), MXD_ENOTSUP, "too many anchors");
  if (valid) {
    if ((rc = check_tensor(valid, "valid", U8, 1, 1, &dev))) return rc;
    MXD_REQUIRE(valid->shape[0] == n, MXD_EINVAL, "valid must be (N)");
  }
  if ((rc = check_tensor(out, "out", U8, 1, 1, &dev))) return rc;
  MXD_REQUIRE(out->shape[0] == n, MXD_EINVAL, "out must be (N)");
  MXD_REQUIRE(((uintptr_t)dptr<float>(anchors) & 15) == 0, MXD_EINVAL, "anchors must be 16-byte aligned");
  if (n == 0) return MXD_OK;
  inside_flags_kernel<<<((int)n + 255) / 256, 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(dptr<float>(anchors)), valid ? dptr<uint8_t>(valid) : nullptr, (int)n,
      (float)img_h, (float)img_w, allowed_border, dptr<uint8_t>(out));
  MXD_POST_LAUNCH("inside_flags");
  return MXD_OK;
}

}  // extern "C"
