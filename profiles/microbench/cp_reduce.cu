// Microbenchmark asked for by the round-1 review: can the TMA engine's reductions
// (cp.reduce.async.bulk.global.shared::cta.add.f32) carry the RoIAlign backward scatter instead of the
// shared-memory tiles?  They bypass the LSU / RED path measured in smem_atomics.cu (~200 G sector-ops/s).
//
// The backward scatters, per (RoI, channel, map row), one segment of ~16 fp32 (64-80 B once padded to 16-byte
// alignment): 4096 RoIs x 256 channels x 16.4 rows = 17.2 M segments per step, into 731 MB of maps (> L2).
// To finish inside the 143 us HBM floor the engine would have to retire 120 G segments/s.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cp_reduce cp_reduce.cu && ./cp_reduce
// Prints, per segment size: bulk reductions/s and GB/s with (a) targets spread over 731 MB (the real case) and
// (b) targets inside a 32 MB L2-resident window, next to per-lane red.global.add.f32 on the same segments.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// Every thread of the CTA issues bulk reductions of seg_bytes from the CTA's shared buffer (all ones).
__global__ void __launch_bounds__(256) bulk_red_kernel(float* dst, uint32_t n_slots, int seg_bytes, int iters) {
  extern __shared__ __align__(128) float buf[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) buf[i] = 1.0f;
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(buf);
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
    const uint32_t slot = hash32(tid * 9781u + it * 7919u + 17u) % n_slots;
    char* g = reinterpret_cast<char*>(dst) + (size_t)slot * 16;
    const uint32_t s = sbase + ((threadIdx.x * 16) & 8191);
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(g), "r"(s),
                 "r"(seg_bytes)
                 : "memory");
    if ((it & 7) == 7) {
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    }
  }
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// The same segments with per-lane RED: a warp covers 32 consecutive floats per instruction.
__global__ void __launch_bounds__(256) lane_red_kernel(float* dst, uint32_t n_slots, int seg_bytes, int iters) {
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int words = seg_bytes / 4;
  for (int it = 0; it < iters * 32; ++it) {      // one segment per warp-iteration: 32x the iterations of the bulk kernel's threads
    const uint32_t slot = hash32((warp * 32 + (it & 31)) * 9781u + (it >> 5) * 7919u + 17u) % n_slots;
    float* g = dst + (size_t)slot * 4;
    for (int w = lane; w < words; w += 32) atomicAdd(g + w, 1.0f);
  }
}

int main() {
  const size_t big = 731ull << 20, small = 32ull << 20;
  float* dst;
  cudaMalloc(&dst, big + 4096);
  cudaMemset(dst, 0, big + 4096);
  cudaFuncSetAttribute(bulk_red_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * 4, block = 256, iters = 64;
  const double n_seg = (double)grid * block * iters;
  printf("%-8s %-10s %14s %10s %14s %10s\n", "seg B", "target", "bulk Gseg/s", "GB/s", "lane-RED Gseg/s", "GB/s");
  const int sizes[] = {16, 32, 64, 128, 256, 1024};
  for (int si = 0; si < 6; ++si) {
    for (int t = 0; t < 2; ++t) {
      const size_t span = t == 0 ? big : small;
      const uint32_t n_slots = (uint32_t)((span - sizes[si]) / 16);
      float ms[2];
      for (int k = 0; k < 2; ++k) {
        for (int rep = 0; rep < 2; ++rep) {           // second repetition is the timed one
          cudaEventRecord(e0);
          if (k == 0) bulk_red_kernel<<<grid, block, 32768>>>(dst, n_slots, sizes[si], iters);
          else lane_red_kernel<<<grid, block>>>(dst, n_slots, sizes[si], iters);
          cudaEventRecord(e1);
          if (cudaEventSynchronize(e1) != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
          cudaEventElapsedTime(&ms[k], e0, e1);
        }
      }
      printf("%-8d %-10s %14.2f %10.1f %14.2f %10.1f\n", sizes[si], t == 0 ? "731 MB" : "32 MB (L2)",
             n_seg / ms[0] / 1e6, n_seg * sizes[si] / ms[0] / 1e6, n_seg / ms[1] / 1e6, n_seg * sizes[si] / ms[1] / 1e6);
    }
  }
  return 0;
}
