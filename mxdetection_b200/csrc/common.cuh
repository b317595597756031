// Shared helpers for libmxdet_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include <math.h>
#include <mutex>
#include "../../include/mxdet.h"

namespace mxd {

// ---- error plumbing ---------------------------------------------------------
int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);

#define MXD_REQUIRE(cond, code, ...)                 \
  do {                                               \
    if (!(cond)) return ::mxd::set_error((code), __VA_ARGS__); \
  } while (0)

#define MXD_CUDA_OK(expr)                                                        \
  do {                                                                           \
    cudaError_t _e = (expr);                                                     \
    if (_e != cudaSuccess)                                                       \
      return ::mxd::set_error(MXD_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

// Call after every kernel launch: counts it and surfaces launch-config errors.
#define MXD_POST_LAUNCH(name)                                                    \
  do {                                                                           \
    ::mxd::count_launch();                                                       \
    cudaError_t _e = cudaPeekAtLastError();                                      \
    if (_e != cudaSuccess) {                                                     \
      cudaGetLastError();                                                        \
      return ::mxd::set_error(MXD_ECUDA, "launch of %s failed: %s", name, cudaGetErrorString(_e)); \
    }                                                                            \
  } while (0)

// ---- DLTensor validation ------------------------------------------------------
enum DT { F32, I32, U8 };

inline bool dtype_is(const DLTensor* t, DT d) {
  if (t->dtype.lanes != 1) return false;
  switch (d) {
    case F32: return t->dtype.code == kDLFloat && t->dtype.bits == 32;
    case I32: return t->dtype.code == kDLInt && t->dtype.bits == 32;
    case U8:  return (t->dtype.code == kDLUInt || t->dtype.code == kDLBool) && t->dtype.bits == 8;
  }
  return false;
}

inline bool is_compact(const DLTensor* t) {
  if (!t->strides) return true;
  for (int i = 0; i < t->ndim; ++i)
    if (t->shape[i] == 0) return true;       // no element: any stride vector describes it (exporters differ)
  int64_t s = 1;
  for (int i = t->ndim - 1; i >= 0; --i) {
    if (t->shape[i] != 1 && t->strides[i] != s) return false;
    s *= t->shape[i];
  }
  return true;
}

inline int64_t numel(const DLTensor* t) {
  int64_t n = 1;
  for (int i = 0; i < t->ndim; ++i) n *= t->shape[i];
  return n;
}

// Validates device / dtype / rank / contiguity; *dev tracks the common device.
int check_tensor(const DLTensor* t, const char* name, DT dt, int ndim_lo, int ndim_hi, int* dev);

template <typename T>
inline T* dptr(const DLTensor* t) {
  return reinterpret_cast<T*>(static_cast<char*>(t->data) + t->byte_offset);
}

inline cudaStream_t as_stream(void* s) { return static_cast<cudaStream_t>(s); }

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Opt-in dynamic shared memory is a per-DEVICE function attribute: call sites set it once per device,
//   static unsigned long long seen = 0;
//   DeviceOnce once(&seen); if (once.first()) MXD_CUDA_OK(cudaFuncSetAttribute(...));
// The guard serialises first use: a second host thread arriving meanwhile waits until the attribute is set (it
// would otherwise launch with the default limit and fail), later calls take one atomic load.
std::mutex& device_once_mutex();
struct DeviceOnce {
  unsigned long long* seen;
  unsigned long long bit;
  bool mine;
  explicit DeviceOnce(unsigned long long* s) : seen(s), bit(0), mine(false) {
    int dev = 0;
    cudaGetDevice(&dev);
    bit = 1ull << (dev & 63);
    if (__atomic_load_n(seen, __ATOMIC_ACQUIRE) & bit) return;
    device_once_mutex().lock();
    if (__atomic_load_n(seen, __ATOMIC_RELAXED) & bit) { device_once_mutex().unlock(); return; }
    mine = true;
  }
  DeviceOnce(const DeviceOnce&) = delete;
  DeviceOnce& operator=(const DeviceOnce&) = delete;
  bool first() const { return mine; }
  ~DeviceOnce() {
    if (mine) {
      __atomic_fetch_or(seen, bit, __ATOMIC_RELEASE);
      device_once_mutex().unlock();
    }
  }
};

// ---- programmatic dependent launch ---------------------------------------------------
// A kernel launched with launch_pdl may be scheduled while its predecessor in the stream is still running (the launch
// latency and the kernel's own prologue overlap the predecessor's tail); it must call pdl_wait() before it touches
// anything the predecessor wrote.  Planner -> persistent-kernel chains use it: 3-4 us per kernel boundary.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_if(bool enable, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = (!enable || getenv("MXD_NO_PDL")) ? 0 : 1;      // (env: A/B switch, plain stream order)
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  return launch_pdl_if(true, kernel, grid, block, smem, st, args...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// ---- device helpers -------------------------------------------------------------
// float -> uint32 whose unsigned order equals the float order (-0 == +0).
__device__ __forceinline__ uint32_t f32_orderable(float x) {
  uint32_t u = __float_as_uint(__fadd_rn(x, 0.0f));
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f32_from_orderable(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

// Spec B / Spec D intersection-over-union pieces, strict fp32, no contraction.
__device__ __forceinline__ float box_area_clamped(float x1, float y1, float x2, float y2, float d) {
  float w = fmaxf(__fadd_rn(__fsub_rn(x2, x1), d), 0.0f);
  float h = fmaxf(__fadd_rn(__fsub_rn(y2, y1), d), 0.0f);
  return __fmul_rn(w, h);
}
__device__ __forceinline__ float box_area_raw(float x1, float y1, float x2, float y2, float d) {
  return __fmul_rn(__fadd_rn(__fsub_rn(x2, x1), d), __fadd_rn(__fsub_rn(y2, y1), d));
}
// Returns true and sets iou when the boxes intersect with positive iw, ih.
__device__ __forceinline__ bool box_iou_pos(const float4& a, float area_a, const float4& b,
                                            float area_b, float d, float* iou) {
  float iw = __fadd_rn(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), d);
  float ih = __fadd_rn(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), d);
  if (!(iw > 0.0f) || !(ih > 0.0f)) return false;
  float inter = __fmul_rn(iw, ih);
  *iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
  return true;
}

// Exactly `box_iou_pos(...) && iou > thr` (Spec B's strict test on the ROUNDED fp32 quotient) without the IEEE
// division for all but borderline pairs: with p = fl(thr*union), inter > p*(1+3e-7) implies the real quotient
// exceeds thr by more than one ulp (so its rounding is > thr) and inter < p*(1-3e-7) implies it is below thr (so
// its rounding is <= thr); only the sliver in between pays for __fdiv_rn.
__device__ __forceinline__ bool box_iou_gt(const float4& a, float area_a, const float4& b, float area_b, float d,
                                           float thr) {
  const float iw = __fadd_rn(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), d);
  const float ih = __fadd_rn(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), d);
  if (!(iw > 0.0f) || !(ih > 0.0f)) return false;
  const float inter = __fmul_rn(iw, ih);
  const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
  // straight-line form of: if (guard) { if (inter > p*(1+e)) return true; if (inter < p*(1-e)) return false; }
  // - same products, same comparisons, one rarely taken branch instead of three
  const float p = __fmul_rn(thr, uni);
  const bool yes = inter > __fmul_rn(p, 1.00000036f);
  const bool no = inter < __fmul_rn(p, 0.99999964f);
  const bool guard = thr > 0.0f && uni > 0.0f && uni < 3.0e38f;
  if (guard && (yes || no)) return yes;
  return __fdiv_rn(inter, uni) > thr;
}

// Correctly rounded fp32 exp/log (fp64 evaluation, one rounding) - Spec F.
__device__ __forceinline__ float exp_cr(float x) { return (float)exp((double)x); }
__device__ __forceinline__ float log_cr(float x) { return (float)log((double)x); }

// Spec G canonical threshold form.
__device__ __forceinline__ int roi_level(float x1, float y1, float x2, float y2, int num_levels,
                                         float finest_scale) {
  float w = __fadd_rn(__fsub_rn(x2, x1), 1.0f), h = __fadd_rn(__fsub_rn(y2, y1), 1.0f);
  float s = __fsqrt_rn(__fmul_rn(w, h));
  float v = __fadd_rn(__fdiv_rn(s, finest_scale), 1e-6f);
  int lvl = 0;
  float t = 2.0f;
  for (int i = 1; i < num_levels; ++i) {
    lvl += (v >= t) ? 1 : 0;
    t = __fmul_rn(t, 2.0f);
  }
  return lvl;
}

}  // namespace mxd
