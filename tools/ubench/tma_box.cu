// Micro-test: where does a multi-dimensional TMA box land in shared memory when its inner dimension (32 bytes) is narrower than
// the 128-byte swizzle span?  (patch embedding: box = (p2 16, y 4, px npx, c 1, b nb) over a [B, C, H, W] bf16 image)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../noise-robust-vit_b200/csrc -L../../noise-robust-vit_b200/lib -lnrvit -o tma_box tma_box.cu
#include "common.cuh"
#include <cstdio>
#include <vector>
using namespace nrv;
__global__ void k(const __grid_constant__ CUtensorMap tm, uint16_t* out, int nbytes, int tx_bytes, int y, int px, int b) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  for (int i = threadIdx.x; i < nbytes / 2; i += blockDim.x) reinterpret_cast<uint16_t*>(smem)[i] = 0xFFFF;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  fence_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(smem_u32(&bar), tx_bytes);
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(&bar)), "r"(0), "r"(y), "r"(px), "r"(0), "r"(b) : "memory");
    mbar_wait(smem_u32(&bar), 0, 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nbytes / 2; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}
int main() {
  if (nrv_init(0)) { printf("init failed: %s\n", nrv_last_error()); return 1; }
  const int B = 3, C = 1, H = 16, W = 64, pw = 16, gw = W / pw;
  std::vector<uint16_t> h(B * C * H * W);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint16_t)i;     // value = b*1024 + y*64 + x
  uint16_t *d, *o;
  cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  const int nbytes = 16384;
  cudaMalloc(&o, nbytes);
  const uint64_t dims[5] = {(uint64_t)pw, (uint64_t)H, (uint64_t)gw, (uint64_t)C, (uint64_t)B};
  const uint64_t strides[4] = {(uint64_t)W * 2, (uint64_t)pw * 2, (uint64_t)H * W * 2, (uint64_t)C * H * W * 2};
  for (int mode = 0; mode < 2; ++mode) {
    const uint32_t box[5] = {(uint32_t)pw, 4, 2, 1, 4};          // 16 x 4 x 2 x 1 x 4 = 512 elements = 1 KB; images 3.. are out of bounds
    CUtensorMap tm;
    if (encode_tmap_5d(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, d, dims, strides, box, mode ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE)) {
      printf("encode failed: %s\n", nrv_last_error()); return 1;
    }
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, nbytes + 1024);
    k<<<1, 128, nbytes + 1024>>>(tm, o, nbytes, 1024, 4, 1, 0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("mode %s: %s\n", mode ? "SWIZZLE_128B" : "no swizzle", cudaGetErrorString(e));
    std::vector<uint16_t> r(nbytes / 2);
    cudaMemcpy(r.data(), o, nbytes, cudaMemcpyDeviceToHost);
    int last = -1;
    for (int c = 0; c < nbytes / 16; ++c) {                        // one line per 16-byte chunk that was written
      const uint16_t v = r[c * 8];
      bool written = false;
      for (int e2 = 0; e2 < 8; ++e2) written = written || r[c * 8 + e2] != 0xFFFF;
      if (!written) continue;
      last = c;
      if (c < 80) printf("  chunk %3d (row128 %2d, unit %d): first = b %d y %2d x %2d  (px %d p2 %2d)%s\n", c, c / 8, c % 8, v / 1024, (v % 1024) / 64,
                         v % 64, (v % 64) / 16, v % 16, v == 0 ? "  [zero fill or element 0]" : "");
    }
    printf("  last written chunk: %d (of %d)\n", last, nbytes / 16);
  }
  return 0;
}
