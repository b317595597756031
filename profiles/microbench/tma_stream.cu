// How fast can one CTA per SM stream a big buffer through a shared-memory ring with cp.async.bulk?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream tma_stream.cu && ./tma_stream
// Per SM: DEPTH copies of BYTES in flight, one thread re-issues a slot as soon as its copy has landed (nobody reads
// the data).  Prints aggregate GB/s for a few (BYTES, DEPTH): the ceiling of any TMA-fed persistent kernel.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128, 1) stream(const char* src, size_t per_cta, int bytes, int depth, int iters) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) u64 bar[32];
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const char* base = src + (size_t)blockIdx.x * per_cta;
  size_t off = 0;
  for (int it = 0; it < iters + depth; ++it) {
    const int s = it % depth;
    if (it >= depth) {
      const uint32_t par = ((it / depth) - 1) & 1;
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(s32(&bar[s])), "r"(par) : "memory");
    }
    if (it < iters) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[s])), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(s32(smem + (size_t)s * bytes)), "l"(base + off), "r"(bytes), "r"(s32(&bar[s])) : "memory");
      off += bytes;
      if (off + bytes > per_cta) off = 0;
    }
  }
}
int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t per_cta = 6u << 20, total = per_cta * sms;     // 888 MB > L2
  char* src;
  cudaMalloc(&src, total);
  cudaMemset(src, 1, total);
  cudaFuncSetAttribute(stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const int sizes[] = {1344, 2688, 5376, 10752, 21504, 43008};
  for (int bytes : sizes)
    for (int depth : {2, 4, 8, 16}) {
      if ((size_t)bytes * depth > 220 * 1024) continue;
      const int iters = (int)(per_cta / bytes);
      stream<<<sms, 128, (size_t)bytes * depth>>>(src, per_cta, bytes, depth, iters);
      cudaEventRecord(a);
      stream<<<sms, 128, (size_t)bytes * depth>>>(src, per_cta, bytes, depth, iters);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms;
      cudaEventElapsedTime(&ms, a, b);
      printf("copy %6d B x depth %2d (%6.1f KB in flight per SM): %7.1f GB/s  %s\n", bytes, depth, bytes * depth / 1024.0,
             (double)iters * bytes * sms / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
