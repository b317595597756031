// Library-level entry points and the validation / error helpers.
#include <atomic>
#include "common.cuh"

namespace mxd {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

std::mutex& device_once_mutex() {
  static std::mutex mu;
  return mu;
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int check_tensor(const DLTensor* t, const char* name, DT dt, int ndim_lo, int ndim_hi, int* dev) {
  MXD_REQUIRE(t != nullptr, MXD_EINVAL, "%s: null tensor", name);
  if (t->device.device_type != kDLCUDA && t->device.device_type != kDLCUDAManaged)
    return set_error(MXD_ENOTSUP, "%s: device_type %d is not CUDA (no CPU fallback in libmxdet_sm100)",
                     name, (int)t->device.device_type);
  MXD_REQUIRE(dtype_is(t, dt), MXD_EINVAL, "%s: dtype code=%d bits=%d lanes=%d, expected %s", name,
              (int)t->dtype.code, (int)t->dtype.bits, (int)t->dtype.lanes,
              dt == F32 ? "float32" : dt == I32 ? "int32" : "uint8");
  MXD_REQUIRE(t->ndim >= ndim_lo && t->ndim <= ndim_hi, MXD_EINVAL, "%s: ndim %d not in [%d,%d]",
              name, t->ndim, ndim_lo, ndim_hi);
  for (int i = 0; i < t->ndim; ++i)
    MXD_REQUIRE(t->shape[i] >= 0, MXD_EINVAL, "%s: negative extent", name);
  MXD_REQUIRE(is_compact(t), MXD_EINVAL, "%s: tensor must be compact row-major", name);
  MXD_REQUIRE(numel(t) == 0 || t->data != nullptr, MXD_EINVAL, "%s: null data pointer", name);
  if (dev) {
    if (*dev < 0) *dev = t->device.device_id;
    MXD_REQUIRE(*dev == t->device.device_id, MXD_EINVAL, "%s: on device %d, other tensors on %d",
                name, t->device.device_id, *dev);
  }
  return MXD_OK;
}

}  // namespace mxd

extern "C" {
int mxd_version(void) { return 100; /* 0.1.0 */ }
const char* mxd_last_error(void) { return mxd::g_err; }
int mxd_sizeof_rpn_config(void) { return (int)sizeof(mxd_rpn_config); }
uint64_t mxd_launch_count(void) { return mxd::g_launches.load(std::memory_order_relaxed); }

// Strided host<->device copy of `rows` runs of `width_bytes` (cudaMemcpy2DAsync): the host-buffer RoI stage moves
// channel slices of (R, C, PH, PW) tensors without first compacting them on the CPU.  kind: 1 = H2D, 2 = D2H.
int mxd_copy2d_async(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes, size_t rows,
                     int kind, void* stream) {
  using namespace mxd;
  MXD_REQUIRE(kind == 1 || kind == 2, MXD_EINVAL, "kind must be 1 (H2D) or 2 (D2H)");
  if (width_bytes == 0 || rows == 0) return MXD_OK;
  MXD_REQUIRE(dst && src && dst_pitch >= width_bytes && src_pitch >= width_bytes, MXD_EINVAL, "bad 2-D copy geometry");
  MXD_CUDA_OK(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width_bytes, rows,
                                kind == 1 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, as_stream(stream)));
  return MXD_OK;
}
}
