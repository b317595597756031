// Why does the tile write-out of the RoIAlign backward run at half of HBM's write rate?  This probe writes zeros to a
// (N, C, H, W) fp32 tensor in exactly the pattern of roi_align_tile_bwd_kernel's write-out - persistent CTAs pull
// (tile, 32-channel slice) items off a counter, warp r of the CTA writes tile row r: for each of the 32 channels the TW
// floats of that row - with no shared memory, no messages, nothing else.  Tile width / height and item order vary.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tile_store tile_store.cu && ./tile_store
#include <cstdio>
#include <cuda_runtime.h>

struct P { int N, C, H, W, th, tw, nty, ntx, order; };

__global__ void __launch_bounds__(1024) tile_store_kernel(float* dst, P p, int* counter) {
  __shared__ int s_item;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = p.nty * p.ntx, ncg = p.C / 32;
  const int n_items = p.N * ncg * tiles;
  const size_t plane = (size_t)p.H * p.W;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_item = atomicAdd(counter, 1);
    __syncthreads();
    const int item = s_item;
    if (item >= n_items) break;
    int tl, r;
    if (p.order == 0) { tl = item % tiles; r = item / tiles; }            // kernel's order: tiles of one (image, slice) adjacent
    else { r = item % (p.N * ncg); tl = item / (p.N * ncg); }             // slices of one tile adjacent
    const int cg = r % ncg, b = r / ncg;
    const int ty0 = (tl / p.ntx) * p.th, tx0 = (tl % p.ntx) * p.tw;
    const int gy = ty0 + warp;
    if (warp >= p.th || gy >= p.H) continue;
    const int twe = min(p.tw, p.W - tx0);
    float* g0 = dst + (((size_t)b * p.C + cg * 32) * p.H + gy) * p.W + tx0;
    for (int x = lane; x < twe; x += 32) {
      float* gp = g0 + x;
#pragma unroll 8
      for (int ch = 0; ch < 32; ++ch) { *gp = 0.0f; gp += plane; }
    }
  }
}

int main() {
  const int N = 8, C = 256, H = 200, W = 336;
  const size_t bytes = (size_t)N * C * H * W * 4;
  float* dst; int* counter;
  cudaMalloc(&dst, bytes); cudaMalloc(&counter, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  cudaEventRecord(e0); cudaMemsetAsync(dst, 0, bytes); cudaEventRecord(e1); cudaEventSynchronize(e1);
  cudaEventRecord(e0); cudaMemsetAsync(dst, 0, bytes); cudaEventRecord(e1); cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms, e0, e1);
  printf("cudaMemset of %.0f MB: %.3f ms = %.0f GB/s\n", bytes / 1e6, ms, bytes / ms / 1e6);
  const int cfgs[][3] = {{25, 48, 0}, {25, 48, 1}, {25, 84, 0}, {25, 112, 0}, {25, 168, 0}, {25, 336, 0}, {12, 336, 0}, {8, 336, 0}, {25, 32, 0}};
  for (auto& c : cfgs) {
    P p{N, C, H, W, c[0], c[1], (H + c[0] - 1) / c[0], (W + c[1] - 1) / c[1], c[2]};
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemsetAsync(counter, 0, 4);
      cudaEventRecord(e0);
      tile_store_kernel<<<148, 32 * c[0]>>>(dst, p, counter);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("tile %2d rows x %3d cols, order %d: %.3f ms = %.0f GB/s (%s)\n", c[0], c[1], c[2], ms, bytes / ms / 1e6,
           cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
