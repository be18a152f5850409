// Throughput of FFMA vs FFMA2 (fma.rn.f32x2) per SM on sm_100a: 16 warps / SM, 16 independent accumulators per thread.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters, float w) {
  float2 acc[16];
  for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  float2 v = make_float2(threadIdx.x * 1e-4f, 0.25f);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) {          // packed: one FFMA2 per two FMAs
        acc[i] = __ffma2_rn(make_float2(w, w), v, acc[i]);
      } else {                  // scalar: two FFMA
        acc[i].x = fmaf(w, v.x, acc[i].x);
        acc[i].y = fmaf(w, v.y, acc[i].y);
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0)
    printf("mode %d: %.3f cycles per (16 warps x 32 FMAs-per-thread step) -> %.1f FMA/clk/SM\n", MODE,
           double(t1 - t0) / iters, 512.0 * 32 * iters / double(t1 - t0));
}
int main() {
  float* out;
  cudaMalloc(&out, 148 * 512 * 4);
  for (int rep = 0; rep < 2; ++rep) {
    k<0><<<148, 512>>>(out, 20000, 1.0001f);
    cudaDeviceSynchronize();
    k<1><<<148, 512>>>(out, 20000, 1.0001f);
    cudaDeviceSynchronize();
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
