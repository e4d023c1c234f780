// Micro-benchmark: tcgen05.mma issue/execute rate for the small-N shapes of the attention kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../noise-robust-vit_b200/csrc -o mma mma.cu
#include "common.cuh"
#include <cstdio>
using namespace nrv;
// mode: 0 SS K/K ; 1 SS K/MN ; 2 TS (A tmem) B MN ; 3 SS MN(2 chunks)/MN
__global__ void __launch_bounds__(128, 1) k(long long* clk, int mode, int N, int count, int reps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sb = smem_u32(smem);
  __shared__ uint32_t tptr;
  __shared__ uint64_t bar;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  fence_async_smem();
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    if (elect_one()) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc(smem_u32(&tptr), 512);
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t T = tptr;
  if (warp == 0 && elect_one()) {
    const uint32_t a_mn = mode == 3, b_mn = mode != 0;
    const uint32_t idesc = make_idesc(1u, a_mn, b_mn, 128u, (uint32_t)N);
    const uint64_t ad = mode == 3 ? make_smem_desc_sw128(sb, 16384, 1024) : make_smem_desc_sw128(sb, 16, 1024);
    const uint64_t bd = make_smem_desc_sw128(sb + 32768, 16, 1024);
    uint32_t ph = 0;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int i = 0; i < count; ++i) {
        const int ks = i & 3;
        const uint64_t a = mode == 3 ? ad + ks * (2048 >> 4) : ad + 2 * ks;
        const uint64_t b = b_mn ? bd + ks * (2048 >> 4) : bd + 2 * ks;
        if (mode == 2) umma_bf16_ts(T + 256, T + ks * 16, b, idesc, 1u);
        else umma_bf16(T + 256, a, b, idesc, 1u);
      }
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), ph, 1);
      ph ^= 1;
    }
    long long t1 = clock64();
    clk[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(T, 512);
}
int main() {
  long long* c; cudaMalloc(&c, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const char* names[4] = {"SS A K-major  B K-major ", "SS A K-major  B MN-major", "TS A in TMEM  B MN-major", "SS A MN-major B MN-major"};
  for (int mode = 0; mode < 4; ++mode)
    for (int N : {16, 64, 128, 256})
      for (int count : {8, 64}) {
        if (mode == 3 && N > 64) continue;
        long long h = 0;
        const int reps = 50;
        k<<<1, 128, 100 * 1024>>>(c, mode, N, count, reps);
        cudaError_t e = cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        printf("%s M=128 N=%3d K=16: batch of %2d MMAs + commit/wait: %.1f clk per MMA (%.0f per batch)\n", names[mode], N, count,
               (double)h / (reps * count), (double)h / reps);
      }
  return 0;
}
