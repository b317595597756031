// Does a predicated-OFF shared-memory load / store still cost LSU issue bandwidth on B200?
// (Question left open by the backward's "predicate the trash-column slots off" experiment, which changed nothing.)
// Every warp runs a loop of 8 x (@p ld.shared, @p st.shared) read-modify-writes on its own shared-memory row; p comes from a
// kernel argument (warp-uniform, unknown at compile time).  Compared: p = true, p = false, and the same loop with the
// memory instructions removed (only the FADDs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pred_lds pred_lds.cu && ./pred_lds
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>   // 0: predicated RMWs, 1: no memory instructions
__global__ void __launch_bounds__(1024) k(int iters, int pred, float* sink) {
  __shared__ float buf[1024 * 9];
  const unsigned base = (unsigned)__cvta_generic_to_shared(buf + threadIdx.x);
  buf[threadIdx.x] = 1.0f;
  for (int j = 1; j < 9; ++j) buf[threadIdx.x + j * 1024] = 0.0f;
  __syncthreads();
  float acc = (float)threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MODE == 0) {
        asm volatile(
            "{ .reg .pred q; .reg .f32 t;\n"
            "setp.ne.s32 q, %2, 0;\n"
            "mov.f32 t, 0f00000000;\n"
            "@q ld.shared.f32 t, [%1];\n"
            "add.f32 %0, %0, t;\n"
            "@q st.shared.f32 [%1], %0;\n }"
            : "+f"(acc)
            : "r"(base + j * 4096), "r"(pred));
      } else {
        asm volatile("add.f32 %0, %0, %1;" : "+f"(acc) : "f"((float)j));
      }
    }
  }
  if (acc == 12345.678f) sink[0] = acc;
}

int main() {
  float* sink;
  cudaMalloc(&sink, 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4096;
  auto run = [&](int mode, int pred) {
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148, 1024>>>(iters, pred, sink); else k<1><<<148, 1024>>>(iters, pred, sink);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
    }
    return ms;
  };
  const double rmw = 148.0 * 32 * iters * 8;      // warp-level read-modify-writes
  const float on = run(0, 1), off = run(0, 0), none = run(1, 0);
  printf("predicate true : %.3f ms  (%.2f cycles per warp RMW per SM at 1.965 GHz)\n", on, on * 1e-3 * 1.965e9 / (rmw / 148));
  printf("predicate false: %.3f ms  (%.2f)\n", off, off * 1e-3 * 1.965e9 / (rmw / 148));
  printf("no LDS / STS   : %.3f ms  (%.2f)\n", none, none * 1e-3 * 1.965e9 / (rmw / 148));
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
