/* ABI-compatible restatement of the DLPack tensor structs (DLPack v0.x / v1.0
 * `DLTensor`, `DLManagedTensor`) so that the library has no build-time
 * dependency.  If the real <dlpack/dlpack.h> was included first its
 * definitions are used instead. */
#ifndef MXDET_DLPACK_H_
#define MXDET_DLPACK_H_
#include <stdint.h>
#ifndef DLPACK_DLPACK_H_
#define DLPACK_DLPACK_H_
#ifdef __cplusplus
extern "C" {
#endif
typedef enum {
  kDLCPU = 1, kDLCUDA = 2, kDLCUDAHost = 3, kDLOpenCL = 4, kDLVulkan = 7, kDLMetal = 8,
  kDLVPI = 9, kDLROCM = 10, kDLROCMHost = 11, kDLExtDev = 12, kDLCUDAManaged = 13
} DLDeviceType;
typedef struct { DLDeviceType device_type; int32_t device_id; } DLDevice;
typedef enum { kDLInt = 0U, kDLUInt = 1U, kDLFloat = 2U, kDLOpaqueHandle = 3U, kDLBfloat = 4U,
               kDLComplex = 5U, kDLBool = 6U } DLDataTypeCode;
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } DLDataType;
typedef struct {
  void* data;
  DLDevice device;
  int32_t ndim;
  DLDataType dtype;
  int64_t* shape;
  int64_t* strides;      /* in elements; NULL = compact row-major */
  uint64_t byte_offset;
} DLTensor;
typedef struct DLManagedTensor {
  DLTensor dl_tensor;
  void* manager_ctx;
  void (*deleter)(struct DLManagedTensor* self);
} DLManagedTensor;
#ifdef __cplusplus
}
#endif
#endif /* DLPACK_DLPACK_H_ */
#endif /* MXDET_DLPACK_H_ */
