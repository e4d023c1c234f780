// Micro-benchmark: one softmax "exp chunk" (8 FFMA2, 16 MUFU.EX2, 8 FADD2, 8 F2FP) per 16 scores, as in attention_fwd2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) { __nv_bfloat162 t = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&t); }
template <int MODE>
__global__ void k(uint32_t* out, const float* in, long long* clk, int iters) {
  float v[8][16];
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int j = 0; j < 16; ++j) v[c][j] = in[(c * 16 + j) * 32 + (threadIdx.x & 31)];
  const uint64_t c2 = f2_pack(in[0], in[0]), n2 = f2_pack(in[1], in[1]);
  uint64_t sum2 = f2_pack(0.f, 0.f), sumb = f2_pack(0.f, 0.f);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        float x0, x1;
        f2_unpack(f2_fma(f2_pack(v[c][j], v[c][j + 1]), c2, n2), x0, x1);
        float e0 = ex2(x0), e1 = ex2(x1);
        if (MODE == 0) sum2 = f2_add(sum2, f2_pack(e0, e1));
        if (MODE == 1) { if (j & 2) sum2 = f2_add(sum2, f2_pack(e0, e1)); else sumb = f2_add(sumb, f2_pack(e0, e1)); }
        if (MODE == 2) { float s0, s1; f2_unpack(sum2, s0, s1); s0 += e0; s1 += e1; sum2 = f2_pack(s0, s1); }
        pk[j >> 1] = pack_bf16(e0, e1);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc ^= pk[j];
      v[c][it & 15] += 1e-6f;
    }
  }
  long long t1 = clock64();
  float s0, s1, s2, s3; f2_unpack(sum2, s0, s1); f2_unpack(sumb, s2, s3);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __float_as_uint(s0 + s1 + s2 + s3);
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
int main() {
  uint32_t* o; long long* c; float* in; cudaMalloc(&o, 1 << 20); cudaMalloc(&c, 8); cudaMalloc(&in, 1 << 20); cudaMemset(in, 0, 1 << 20);
  const int iters = 200;
  for (int warps : {4, 8}) {
    long long h;
    k<0><<<1, warps * 32>>>(o, in, c, iters); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("warps/SM %d  chunk(16 exp) one FADD2 chain : %.1f clk per chunk per warp-slot\n", warps, (double)h / (iters * 8.0 * (warps / 4)));
    k<1><<<1, warps * 32>>>(o, in, c, iters); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("warps/SM %d  chunk(16 exp) two FADD2 chains: %.1f\n", warps, (double)h / (iters * 8.0 * (warps / 4)));
    k<2><<<1, warps * 32>>>(o, in, c, iters); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("warps/SM %d  chunk(16 exp) scalar FADD     : %.1f\n", warps, (double)h / (iters * 8.0 * (warps / 4)));
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
