// Micro-benchmark: MUFU.EX2 / FFMA2 issue rates per SM sub-partition on this part.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu mufu.cu && ./mufu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(float* out, long long* clk, int iters) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = -0.001f * (threadIdx.x + i);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) a[i] = ex2(a[i]);
      if (MODE == 1) a[i] = fmaf(a[i], 0.999f, 0.001f);
      if (MODE == 2) { a[i] = ex2(a[i]); a[i] = fmaf(a[i], 0.999f, -0.5f); a[i] = fmaf(a[i], 0.999f, -0.5f); }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
int main() {
  float* o; long long* c; cudaMalloc(&o, 1 << 20); cudaMalloc(&c, 8);
  const int iters = 1000;
  for (int warps : {4, 8, 16}) {
    long long h;
    k<0><<<1, warps * 32>>>(o, c, iters); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("warps/SM %2d  EX2 : %.2f clk per warp-instr per SMSP\n", warps, (double)h / (iters * 16.0 * (warps / 4)));
    k<1><<<1, warps * 32>>>(o, c, iters); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("warps/SM %2d  FFMA: %.2f clk per warp-instr per SMSP\n", warps, (double)h / (iters * 16.0 * (warps / 4)));
    k<2><<<1, warps * 32>>>(o, c, iters); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("warps/SM %2d  EX2+2FFMA: %.2f clk per triple per SMSP\n", warps, (double)h / (iters * 16.0 * (warps / 4)));
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
