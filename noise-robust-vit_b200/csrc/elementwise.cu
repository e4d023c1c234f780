// HBM-bound kernels of the ViT hot path: LayerNorm fwd/bwd, column reductions (bias / affine
// gradients), patch extraction, class-token / positional-embedding glue, pooling, softmax
// cross-entropy, fused AdamW.  All accesses are 16-byte vectorised and coalesced; grids are
// sized from the SM count.  Every kernel is instantiated for bf16 (production) and fp32 (check
// mode) activations.  Reference ops replaced are cited per kernel.
#include "common.cuh"
#include "nrvit_internal.h"

#include <stdlib.h>

namespace nrv {

// ----------------------------------------------------------------------------------------------
// LayerNorm forward: one warp per row, row held in registers (dim <= 256*NCH)
// (aten::native_layer_norm; simple_vit.py:38,54,136 ; vit.py:104,115,167)
// ----------------------------------------------------------------------------------------------
template <typename T, int NCH>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const T* __restrict__ x,
                                                      const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, float eps,
                                                      T* __restrict__ y, float* __restrict__ mean,
                                                      float* __restrict__ rstd, long long rows,
                                                      int dim) {
  pdl_launch_dependents();   // the next kernel in the stream (a PDL-launched GEMM) may start its prologue
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* xr = x + row * dim;
  float v[NCH][8];
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col = c * 256 + lane * 8;
    if (col < dim) {
      V8<T>::load(xr + col, v[c]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[c][j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[c][j] = 0.f;
    }
  }
  const float mu = warp_sum(s) / (float)dim;
  float sq = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col = c * 256 + lane * 8;
    if (col < dim) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[c][j] - mu; sq += d * d; }
    }
  }
  const float var = warp_sum(sq) / (float)dim;
  const float rs = rsqrtf(var + eps);
  if (lane == 0) {
    if (mean) mean[row] = mu;
    if (rstd) rstd[row] = rs;
  }
  T* yr = y + row * dim;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col = c * 256 + lane * 8;
    if (col < dim) {
      float g[8], b[8], o[8];
      V8<float>::load(gamma + col, g);
      V8<float>::load(beta + col, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[c][j] - mu) * rs * g[j] + b[j];
      V8<T>::store(yr + col, o);
    }
  }
}

// bf16 pair -> packed fp32 pair (two ALU ops), and back
__device__ __forceinline__ uint64_t bf2_to_f2(uint32_t u) {
  return f2_pack(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ uint32_t f2_to_bf2(uint64_t v) {
  float a, b;
  f2_unpack(v, a, b);
  return pack_bf16(a, b);
}

// bf16 production forward for dim = NCH * 256: the generic kernel above spends ~16 instructions per element
// (issue-bound at 4.4 TB/s); this one does the row in packed f32x2 arithmetic (~5 per element), keeps gamma / beta
// in registers and walks rows with a grid stride.
template <int NCH>
__global__ void __launch_bounds__(256, 4) ln_fwd_bf16_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float eps,
                                                           bf16* __restrict__ y, float* __restrict__ mean,
                                                           float* __restrict__ rstd, long long rows) {
  pdl_launch_dependents();   // the next kernel in the stream (a PDL-launched GEMM) may start its prologue
  constexpr int DIM = NCH * 256;
  const int lane = threadIdx.x & 31;
  const long long w0 = (long long)blockIdx.x * 8 + (threadIdx.x >> 5), wstride = (long long)gridDim.x * 8;
  // gamma / beta live in shared memory, permuted so that a lane's two LDS.128 per 256-column chunk are contiguous across the
  // warp ([chunk][half][lane][4]): keeping them in registers (48) held the kernel at two CTAs per SM
  __shared__ __align__(16) float gb_s[2][DIM];
  for (int i = threadIdx.x; i < DIM; i += blockDim.x) {
    const int c = i >> 8, l = (i & 255) >> 3, e = i & 7;
    const int j = c * 256 + (e >> 2) * 128 + l * 4 + (e & 3);
    gb_s[0][j] = gamma[i];
    gb_s[1][j] = beta[i];
  }
  __syncthreads();
  const uint32_t gs = smem_u32(&gb_s[0][0]) + lane * 16, bs = smem_u32(&gb_s[1][0]) + lane * 16;
  const float inv_dim = 1.f / (float)DIM;
  // the next row of this warp is requested before the current one is reduced: two rows in flight per warp (with one,
  // half of the stall samples sat on the first use of the loaded row)
  uint4 nq[NCH];
  if (w0 < rows) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + w0 * DIM) + lane;
#pragma unroll
    for (int c = 0; c < NCH; ++c) nq[c] = xr[c * 32];
  }
  for (long long row = w0; row < rows; row += wstride) {
    uint4 cq[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) cq[c] = nq[c];
    if (row + wstride < rows) {
      const uint4* xn = reinterpret_cast<const uint4*>(x + (row + wstride) * DIM) + lane;
#pragma unroll
      for (int c = 0; c < NCH; ++c) nq[c] = xn[c * 32];
    }
    uint64_t v[NCH][4];
    uint64_t s2 = f2_pack(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const uint4 q = cq[c];
      v[c][0] = bf2_to_f2(q.x); v[c][1] = bf2_to_f2(q.y); v[c][2] = bf2_to_f2(q.z); v[c][3] = bf2_to_f2(q.w);
#pragma unroll
      for (int e = 0; e < 4; ++e) s2 = f2_add(s2, v[c][e]);
    }
    float sa, sb;
    f2_unpack(s2, sa, sb);
    const float mu = warp_sum(sa + sb) * inv_dim;
    const uint64_t nmu2 = f2_pack(-mu, -mu);
    uint64_t q2 = f2_pack(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[c][e] = f2_add(v[c][e], nmu2);
        q2 = f2_fma(v[c][e], v[c][e], q2);
      }
    f2_unpack(q2, sa, sb);
    const float rs = rsqrtf(warp_sum(sa + sb) * inv_dim + eps);
    if (lane == 0) {
      if (mean) mean[row] = mu;
      if (rstd) rstd[row] = rs;
    }
    const uint64_t rs2 = f2_pack(rs, rs);
    uint4* yr = reinterpret_cast<uint4*>(y + row * DIM) + lane;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      uint32_t o[4];
#pragma unroll
      const float4 g0 = lds128f(gs + c * 1024), g1 = lds128f(gs + c * 1024 + 512);
      const float4 b0 = lds128f(bs + c * 1024), b1 = lds128f(bs + c * 1024 + 512);
      const uint64_t g2[4] = {f2_pack(g0.x, g0.y), f2_pack(g0.z, g0.w), f2_pack(g1.x, g1.y), f2_pack(g1.z, g1.w)};
      const uint64_t b2[4] = {f2_pack(b0.x, b0.y), f2_pack(b0.z, b0.w), f2_pack(b1.x, b1.y), f2_pack(b1.z, b1.w)};
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = f2_to_bf2(f2_fma(f2_mul(v[c][e], rs2), g2[e], b2[e]));
      yr[c * 32] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// ----------------------------------------------------------------------------------------------
// LayerNorm backward + residual-gradient add + column partials (dgamma, dbeta, colsum(dx)).
// (aten::native_layer_norm_backward + aten::add of the skip connection)
// A row is covered by NW = ceil(dim/256) warps (8 columns per thread), so the per-thread state is
// small (three 8-wide column accumulators) and 24 warps/SM stay resident: the kernel is bound by
// HBM (reads x, dy, dres; writes dx), not by occupancy.  LNB_WARPS/NW rows are in flight per CTA;
// the two row statistics are combined across the NW warps through shared memory.
// Persistent CTAs walk the rows with a grid stride; block partials go to
// workspace[block][3][dim] and are reduced in a fixed order by colreduce_finalize.
// ----------------------------------------------------------------------------------------------
constexpr int LNB_WARPS = 12;

template <typename T, int NW>
__global__ void __launch_bounds__(LNB_WARPS * 32, 2) ln_bwd_kernel(
    const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ mean,
    const float* __restrict__ rstd, const float* __restrict__ gamma, const T* __restrict__ dres,
    T* __restrict__ dx, float* __restrict__ partial, long long rows, int dim, const float* __restrict__ beta,
    T* __restrict__ xn_out) {
  pdl_launch_dependents();   // the next kernel in the stream (a PDL-launched GEMM) may start its prologue
  constexpr int RPB = LNB_WARPS / NW;            // rows in flight per CTA
  extern __shared__ float red[];                 // [3][dim] block partials
  __shared__ float2 stat[2][RPB][NW];            // (s1, s2) partials, double buffered by iteration parity
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int slot = warp / NW;                    // which of the RPB rows this warp works on
  const int wc = warp - slot * NW;               // column group of this warp
  const bool active = slot < RPB;
  for (int i = threadIdx.x; i < 3 * dim; i += blockDim.x) red[i] = 0.f;
  __syncthreads();

  const int col = wc * 256 + lane * 8;
  const bool col_ok = active && col < dim;
  float gam[8], acc_g[8], acc_b[8], acc_c[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc_g[j] = 0.f; acc_b[j] = 0.f; acc_c[j] = 0.f; gam[j] = 0.f; }
  if (col_ok) V8<float>::load(gamma + col, gam);
  const float inv_dim = 1.f / (float)dim;
  const long long rstride = (long long)gridDim.x * RPB;
  const long long iters = (rows + rstride - 1) / rstride;
  for (long long itn = 0; itn < iters; ++itn) {
    const long long row = itn * rstride + (long long)blockIdx.x * RPB + slot;
    const bool row_ok = active && row < rows;
    float xh[8], g[8];
    float s1 = 0.f, s2 = 0.f, rs = 0.f;
    if (row_ok && col_ok) {
      const float mu = mean[row];
      rs = rstd[row];
      float xv[8], dv[8];
      V8<T>::load(x + row * dim + col, xv);
      V8<T>::load(dy + row * dim + col, dv);
      if (xn_out != nullptr) {   // the normalised rows the forward pass did not keep (LayerNorm folded into its GEMM)
        float bt[8], yn[8];
        V8<float>::load(beta + col, bt);
#pragma unroll
        for (int j = 0; j < 8; ++j) yn[j] = fmaf((xv[j] - mu) * rs, gam[j], bt[j]);
        V8<T>::store(xn_out + row * dim + col, yn);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xh[j] = (xv[j] - mu) * rs;
        g[j] = dv[j] * gam[j];
        s1 += g[j];
        s2 += g[j] * xh[j];
        acc_g[j] += dv[j] * xh[j];
        acc_b[j] += dv[j];
      }
    }
    float c1, c2;
    if (NW == 1) {
      c1 = warp_sum(s1) * inv_dim;
      c2 = warp_sum(s2) * inv_dim;
    } else {
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      const int par = (int)(itn & 1);
      if (active && lane == 0) stat[par][slot][wc] = make_float2(s1, s2);
      // the NW warps of this row meet on a named barrier (ids 1..RPB); idle warps skip it
      if (active) {
#define NRV_BAR(ID) asm volatile("bar.sync " #ID ", %0;" ::"r"(NW * 32) : "memory"); break
        switch (slot) {
          case 0: NRV_BAR(1); case 1: NRV_BAR(2); case 2: NRV_BAR(3); case 3: NRV_BAR(4);
          case 4: NRV_BAR(5); case 5: NRV_BAR(6); case 6: NRV_BAR(7); case 7: NRV_BAR(8);
          case 8: NRV_BAR(9); case 9: NRV_BAR(10); case 10: NRV_BAR(11); default: NRV_BAR(12);
        }
#undef NRV_BAR
      }
      float t1 = 0.f, t2 = 0.f;
      if (active) {
#pragma unroll
        for (int w = 0; w < NW; ++w) { const float2 v = stat[par][slot][w]; t1 += v.x; t2 += v.y; }
      }
      c1 = t1 * inv_dim;
      c2 = t2 * inv_dim;
    }
    if (row_ok && col_ok) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = rs * (g[j] - c1 - xh[j] * c2);
      if (dres != nullptr) {
        float r[8];
        V8<T>::load(dres + row * dim + col, r);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += r[j];
      }
      V8<T>::store(dx + row * dim + col, o);
      // column sums of what was actually stored (rounded), so that the bias gradient equals the
      // column sum of the tensor the dW GEMM consumes
#pragma unroll
      for (int j = 0; j < 8; ++j) acc_c[j] += V8<T>::round(o[j]);
    }
  }
  // block reduce through shared-memory atomics (once per kernel, negligible)
  if (col_ok) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&red[col + j], acc_g[j]);
      atomicAdd(&red[dim + col + j], acc_b[j]);
      atomicAdd(&red[2 * dim + col + j], acc_c[j]);
    }
  }
  __syncthreads();
  float* out = partial + (long long)blockIdx.x * 3 * dim;
  for (int i = threadIdx.x; i < 3 * dim; i += blockDim.x) out[i] = red[i];
}

// ----------------------------------------------------------------------------------------------
// LayerNorm backward, bulk-copy staged (bf16 production path, dim = CPL * 256).
// The register kernel above keeps only 2-3 16-byte loads per thread in flight and meets a row barrier every
// iteration, which caps it near 3 TB/s.  Here each CTA walks chunks of 8 consecutive rows; one thread requests the
// three operand tiles of a chunk (x, dy, residual gradient: 8 rows x dim, contiguous in memory) with
// cp.async.bulk into a 3-stage shared-memory ring, so ~70 KB per CTA are in flight with no registers tied up, and
// each warp then owns one whole row: both row sums by shuffles only, no cross-warp barrier.
// Same outputs as ln_bwd_kernel, including the per-CTA partial column sums for colreduce_finalize.
// ----------------------------------------------------------------------------------------------
constexpr int LNT_ROWS = 8, LNT_STAGES = 3;

template <int CPL, bool XN>
__global__ void __launch_bounds__(LNT_ROWS * 32, 2) ln_bwd_tma_kernel(
    const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ mean,
    const float* __restrict__ rstd, const float* __restrict__ gamma, const bf16* __restrict__ dres,
    bf16* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ colsum,
    long long rows, const float* __restrict__ beta, bf16* __restrict__ xn_out) {
  pdl_launch_dependents();   // the next kernel in the stream (a PDL-launched GEMM) may start its prologue
  constexpr int DIM = CPL * 256;
  constexpr int TILE = LNT_ROWS * DIM * 2;          // bytes of one operand tile
  constexpr int STAGE = 3 * TILE + 64;              // x | dy | dres | mean[8] rstd[8] of the chunk's rows
  extern __shared__ __align__(128) uint8_t lsm[];
  float* gam_s = reinterpret_cast<float*>(lsm + LNT_STAGES * STAGE);
  float* bet_s = gam_s + DIM;                       // XN only: beta, permuted like gamma
  const uint32_t sbase = smem_u32(lsm);
  const uint32_t bar0 = sbase + LNT_STAGES * STAGE + DIM * 4 * (XN ? 2 : 1);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // gamma is kept permuted so that the two LDS.128 of a lane (its 8 columns of a 256-column chunk) are each contiguous
  // across the warp: [chunk][half][lane][4]
  for (int i = threadIdx.x; i < DIM; i += blockDim.x) {
    const int c = i >> 8, l = (i & 255) >> 3, e = i & 7;
    gam_s[c * 256 + (e >> 2) * 128 + l * 4 + (e & 3)] = gamma[i];
    if (XN) bet_s[c * 256 + (e >> 2) * 128 + l * 4 + (e & 3)] = beta[i];
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < LNT_STAGES; ++s) mbar_init(bar0 + 8 * s, 1);
    fence_barrier_init();
  }
  __syncthreads();
  const long long chunks = (rows + LNT_ROWS - 1) / LNT_ROWS;
  const long long my_chunks = (long long)blockIdx.x < chunks ? (chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto issue = [&](long long k) {                   // thread 0: request chunk k of this CTA into stage k % STAGES
    const long long row0 = ((long long)blockIdx.x + k * gridDim.x) * LNT_ROWS;
    const long long nrows = rows - row0 < LNT_ROWS ? rows - row0 : LNT_ROWS;
    const uint32_t bytes = (uint32_t)nrows * DIM * 2;
    const int s = (int)(k % LNT_STAGES);
    const uint32_t dst = sbase + s * STAGE, bar = bar0 + 8 * s;
    // full chunks also fetch their rows' statistics (2 x 32 bytes); a ragged last chunk reads them with plain loads
    const uint32_t stat_bytes = nrows == LNT_ROWS ? 2u * LNT_ROWS * 4u : 0u;
    mbar_arrive_expect_tx(bar, bytes * (dres != nullptr ? 3 : 2) + stat_bytes);
    bulk_load_1d(dst, x + row0 * DIM, bytes, bar);
    bulk_load_1d(dst + TILE, dy + row0 * DIM, bytes, bar);
    if (dres != nullptr) bulk_load_1d(dst + 2 * TILE, dres + row0 * DIM, bytes, bar);
    if (stat_bytes) {
      bulk_load_1d(dst + 3 * TILE, mean + row0, LNT_ROWS * 4, bar);
      bulk_load_1d(dst + 3 * TILE + LNT_ROWS * 4, rstd + row0, LNT_ROWS * 4, bar);
    }
  };
  if (threadIdx.x == 0)
    for (long long k = 0; k < LNT_STAGES - 1 && k < my_chunks; ++k) issue(k);

  uint64_t acc_g[CPL][4], acc_b[CPL][4], acc_c[CPL][4];   // packed pairs of column accumulators
#pragma unroll
  for (int k = 0; k < CPL; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc_g[k][j] = f2_pack(0.f, 0.f); acc_b[k][j] = f2_pack(0.f, 0.f); acc_c[k][j] = f2_pack(0.f, 0.f); }
  const float inv_dim = 1.f / (float)DIM;
  for (long long k = 0; k < my_chunks; ++k) {
    // stage (k + STAGES - 1) % STAGES was consumed in iteration k - 1 (closed by the __syncthreads below)
    if (threadIdx.x == 0 && k + LNT_STAGES - 1 < my_chunks) issue(k + LNT_STAGES - 1);
    const long long row = ((long long)blockIdx.x + k * gridDim.x) * LNT_ROWS + warp;
    const bool row_ok = row < rows;
    const int s = (int)(k % LNT_STAGES);
    mbar_wait(bar0 + 8 * s, (uint32_t)((k / LNT_STAGES) & 1), 50);
    // the row's statistics came with the chunk (an LDG one iteration ahead was spilled by ptxas and stalled on its own
    // load: 17 % of the stall samples); only a ragged last chunk reads them from global memory
    float mu = 0.f, rs = 0.f;
    {
      const long long row0 = ((long long)blockIdx.x + k * gridDim.x) * LNT_ROWS;
      if (rows - row0 >= LNT_ROWS) {
        const uint32_t st = sbase + s * STAGE + 3 * TILE + warp * 4;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(mu) : "r"(st) : "memory");
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(rs) : "r"(st + LNT_ROWS * 4) : "memory");
      } else if (row_ok) {
        mu = __ldg(mean + row);
        rs = __ldg(rstd + row);
      }
    }
    if (row_ok) {
      const uint32_t xs = sbase + s * STAGE + warp * DIM * 2 + lane * 16, ds = xs + TILE, rsm = xs + 2 * TILE;
      const uint32_t gs = smem_u32(gam_s) + lane * 16;
      const uint64_t rs2 = f2_pack(rs, rs), nmr2 = f2_pack(-mu * rs, -mu * rs);
      // pass 1: the two row sums  s1 = sum dy*gamma ,  s2 = sum dy*gamma*xhat   (packed f32x2 throughout)
      uint64_t s1 = f2_pack(0.f, 0.f), s2 = f2_pack(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        const uint4 xq = lds128(xs + c * 512), dq = lds128(ds + c * 512);
        const float4 g0 = lds128f(gs + c * 1024), g1 = lds128f(gs + c * 1024 + 512);
        const uint64_t gm[4] = {f2_pack(g0.x, g0.y), f2_pack(g0.z, g0.w), f2_pack(g1.x, g1.y), f2_pack(g1.z, g1.w)};
        const uint32_t xw[4] = {xq.x, xq.y, xq.z, xq.w}, dw[4] = {dq.x, dq.y, dq.z, dq.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint64_t g2 = f2_mul(bf2_to_f2(dw[e]), gm[e]);
          s1 = f2_add(s1, g2);
          s2 = f2_fma(g2, f2_fma(bf2_to_f2(xw[e]), rs2, nmr2), s2);
        }
      }
      float a0, a1, b0, b1;
      f2_unpack(s1, a0, a1);
      f2_unpack(s2, b0, b1);
      const float c1 = warp_sum(a0 + a1) * inv_dim, c2 = warp_sum(b0 + b1) * inv_dim;
      const uint64_t nc1 = f2_pack(-c1, -c1), nc2 = f2_pack(-c2, -c2);
      // pass 2: dx = rs * (dy*gamma - c1 - xhat*c2) + dres, and the column sums (d gamma, d beta, sum of stored dx)
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        const uint4 xq = lds128(xs + c * 512), dq = lds128(ds + c * 512);
        uint4 rq = make_uint4(0, 0, 0, 0);
        if (dres != nullptr) rq = lds128(rsm + c * 512);
        const float4 g0 = lds128f(gs + c * 1024), g1 = lds128f(gs + c * 1024 + 512);
        const uint64_t gm[4] = {f2_pack(g0.x, g0.y), f2_pack(g0.z, g0.w), f2_pack(g1.x, g1.y), f2_pack(g1.z, g1.w)};
        const uint32_t xw[4] = {xq.x, xq.y, xq.z, xq.w}, dw[4] = {dq.x, dq.y, dq.z, dq.w}, rw[4] = {rq.x, rq.y, rq.z, rq.w};
        uint32_t ow[4];
        if (XN) {
          // the normalised row the forward pass did not keep (its LayerNorm is folded into the projection GEMM): the B
          // operand of the weight-gradient GEMM that follows
          const float4 b0 = lds128f(smem_u32(bet_s) + lane * 16 + c * 1024), b1 = lds128f(smem_u32(bet_s) + lane * 16 + c * 1024 + 512);
          const uint64_t bt[4] = {f2_pack(b0.x, b0.y), f2_pack(b0.z, b0.w), f2_pack(b1.x, b1.y), f2_pack(b1.z, b1.w)};
          uint32_t nw[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) nw[e] = f2_to_bf2(f2_fma(f2_fma(bf2_to_f2(xw[e]), rs2, nmr2), gm[e], bt[e]));
          *reinterpret_cast<uint4*>(xn_out + row * DIM + c * 256 + lane * 8) = make_uint4(nw[0], nw[1], nw[2], nw[3]);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint64_t dv = bf2_to_f2(dw[e]);
          const uint64_t xh = f2_fma(bf2_to_f2(xw[e]), rs2, nmr2);
          const uint64_t t = f2_add(f2_fma(xh, nc2, f2_mul(dv, gm[e])), nc1);
          ow[e] = f2_to_bf2(f2_fma(t, rs2, bf2_to_f2(rw[e])));
          acc_g[c][e] = f2_fma(dv, xh, acc_g[c][e]);
          acc_b[c][e] = f2_add(acc_b[c][e], dv);
          acc_c[c][e] = f2_add(acc_c[c][e], bf2_to_f2(ow[e]));   // what the dW GEMM will read back
        }
        *reinterpret_cast<uint4*>(dx + row * DIM + c * 256 + lane * 8) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      }
    }
    __syncthreads();
  }
  // block reduce: the operand ring is idle now, reuse it as [8 warps][3][DIM] floats (72 KB at dim 768).  Every warp writes
  // its 72 packed accumulators with 16-byte stores, then each thread sums four columns over the eight warps.  (Shared-memory
  // float atomics are compare-and-swap loops: with eight warps on the same 2 304 addresses they were 24 % of this kernel's
  // stall samples, ncu source view of round 2.)
  float* red = reinterpret_cast<float*>(lsm);
  {
    float* mine = red + warp * 3 * DIM;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      float v[3][8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        f2_unpack(acc_g[c][j], v[0][2 * j], v[0][2 * j + 1]);
        f2_unpack(acc_b[c][j], v[1][2 * j], v[1][2 * j + 1]);
        f2_unpack(acc_c[c][j], v[2][2 * j], v[2][2 * j + 1]);
      }
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        float4* dst = reinterpret_cast<float4*>(mine + a * DIM + c * 256 + lane * 8);
        dst[0] = make_float4(v[a][0], v[a][1], v[a][2], v[a][3]);
        dst[1] = make_float4(v[a][4], v[a][5], v[a][6], v[a][7]);
      }
    }
  }
  __syncthreads();
  // the CTAs' column sums meet in fp32 vector reds on the outputs (+=): no partial buffer, no finalize launch
  for (int i = threadIdx.x * 4; i < 3 * DIM; i += blockDim.x * 4) {
    float4 t = *reinterpret_cast<const float4*>(red + i);
#pragma unroll
    for (int w = 1; w < LNT_ROWS; ++w) {
      const float4 u = *reinterpret_cast<const float4*>(red + w * 3 * DIM + i);
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    const int k = i / DIM, c = i - k * DIM;
    float* o = k == 0 ? dgamma : (k == 1 ? dbeta : colsum);
    if (o != nullptr)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + c), "f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w) : "memory");
  }
}

template <int CPL, bool XN>
static int launch_ln_bwd_tma(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                             const void* dres, void* dx, float* dgamma, float* dbeta, float* colsum, long long rows, int blocks,
                             cudaStream_t st, const float* beta, void* xn_out) {
  constexpr int DIM = CPL * 256;
  const int smem = LNT_STAGES * (3 * LNT_ROWS * DIM * 2 + 64) + DIM * 4 * (XN ? 2 : 1) + LNT_STAGES * 8 + 16;
  static bool attr_set = false;
  if (!attr_set) {
    NRV_CUDA(cudaFuncSetAttribute(ln_bwd_tma_kernel<CPL, XN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  ln_bwd_tma_kernel<CPL, XN><<<blocks, LNT_ROWS * 32, smem, st>>>((const bf16*)dy, (const bf16*)x, mean, rstd, gamma,
                                                                 (const bf16*)dres, (bf16*)dx, dgamma, dbeta, colsum, rows,
                                                                 beta, (bf16*)xn_out);
  return NRV_OK;
}

// out_k[c] += sum_p partial[p][k][c]   (k < nk; out_k may be NULL).  32 columns x 8 part-lanes per CTA.
__global__ void __launch_bounds__(256) colreduce_finalize(const float* __restrict__ partial, int nparts, int nk,
                                                          int ncols, float* o0, float* o1, float* o2) {
  __shared__ float sm[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int idx = blockIdx.x * 32 + cx;   // flat (k, col)
  float s = 0.f;
  if (idx < nk * ncols)
    for (int p = ry; p < nparts; p += 8) s += partial[(long long)p * nk * ncols + idx];
  sm[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && idx < nk * ncols) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += sm[r][cx];
    const int k = idx / ncols, c = idx - k * ncols;
    float* o = k == 0 ? o0 : (k == 1 ? o1 : o2);
    if (o != nullptr) o[c] += t;
  }
}

// ----------------------------------------------------------------------------------------------
// column sums of a matrix: bias gradients of nn.Linear (autograd of addmm's bias)
// grid = (col strips of 256, row splits); partial[split][cols]
// ----------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, long long ldx,
                                                      long long rows, int cols, int period, int skip,
                                                      float* __restrict__ partial) {
  __shared__ float red[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + lane * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (col < cols) {
    for (long long r = (long long)blockIdx.y * 8 + warp; r < rows; r += (long long)gridDim.y * 8) {
      if (skip > 0 && (int)(r % period) < skip) continue;  // e.g. class-token rows
      float v[8];
      V8<T>::load(x + r * ldx + col, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = acc[j];
  __syncthreads();
  const int t = threadIdx.x;
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w][t];
  const int c = blockIdx.x * 256 + t;
  if (c < cols) partial[(long long)blockIdx.y * cols + c] = s;
}

// Single-pass variant: a warp owns one 256-column strip and a row-interleaved slice (four 16-byte loads per lane in
// flight); the slices meet in fp32 vector reds on `out`, so there are no partials and no finalize launch.
// (Measured and dropped: running it on a side stream beside the persistent GEMMs.  Its CTAs are not scheduled
// next to a 225 KB / 168-register GEMM CTA, so it only delayed the next GEMM on the main stream.)
template <typename T>
__global__ void __launch_bounds__(256, 4) colsum_atomic_kernel(const T* __restrict__ x, long long ldx, long long rows,
                                                           int cols, float* __restrict__ out) {
  pdl_launch_dependents();   // the next kernel in the stream (a PDL-launched GEMM) may start its prologue
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int strips = (cols + 255) >> 8;
  const int wps = (gridDim.x * 8) / strips;           // warps per strip
  const int strip = gw % strips, slot = gw / strips;
  const int col = strip * 256 + lane * 8;
  if (slot >= wps || col >= cols) return;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const T* px = x + col;
  long long r = slot;
  for (; r + 3LL * wps < rows; r += 4LL * wps) {
    float v0[8], v1[8], v2[8], v3[8];
    V8<T>::load(px + r * ldx, v0);
    V8<T>::load(px + (r + wps) * ldx, v1);
    V8<T>::load(px + (r + 2LL * wps) * ldx, v2);
    V8<T>::load(px + (r + 3LL * wps) * ldx, v3);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += (v0[j] + v1[j]) + (v2[j] + v3[j]);
  }
  for (; r < rows; r += wps) {
    float v[8];
    V8<T>::load(px + r * ldx, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += v[j];
  }
  float* o = out + col;
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(acc[0]), "f"(acc[1]), "f"(acc[2]), "f"(acc[3]) : "memory");
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4), "f"(acc[4]), "f"(acc[5]), "f"(acc[6]), "f"(acc[7]) : "memory");
}

// ----------------------------------------------------------------------------------------------
// Dropout (nn.Dropout inside the encoder: vit.py:45,47,109,125,166 ; README ViT emb_dropout / dropout).
// Counter-based: the keep decision of element i at (layer, site) is a pure function of (seed, layer, site, i)
// -- Philox4x32-10, four elements per call -- so backward regenerates the mask instead of storing it.
//   out[i] = x[i] * keep_i / (1 - p)  (+ residual[i])
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
  }
  return c;
}

template <typename T>
__global__ void __launch_bounds__(256) dropout_kernel(const T* __restrict__ x, const T* __restrict__ residual,
                                                      T* __restrict__ out, long long groups, float p, float scale,
                                                      uint2 key, uint32_t stream_id) {
  pdl_launch_dependents();   // the next kernel in the stream (a PDL-launched GEMM) may start its prologue
  // keep iff the top 24 random bits, as a uniform in [0,1), are >= p
  const uint32_t thresh = (uint32_t)(p * 16777216.0f);
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (long long)gridDim.x * blockDim.x) {
    float v[8], r[8];
    V8<T>::load(x + g * 8, v);
    if (residual != nullptr) V8<T>::load(residual + g * 8, r);
    const unsigned long long q = (unsigned long long)g * 2;
    const uint4 a = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)(q >> 32), stream_id, 0u), key);
    const uint4 b = philox4x32_10(make_uint4((uint32_t)(q + 1), (uint32_t)((q + 1) >> 32), stream_id, 0u), key);
    const uint32_t bits[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float m = (bits[j] >> 8) >= thresh ? scale : 0.f;
      v[j] = residual != nullptr ? fmaf(v[j], m, r[j]) : v[j] * m;
    }
    V8<T>::store(out + g * 8, v);
  }
}

// ----------------------------------------------------------------------------------------------
// Noisy-input objective of the examples: x + std * randn_like(x)   (examples/nowak.py:153, default std :196).
// One pass: Philox4x32-10 -> Box-Muller (two normals per pair of uniforms, MUFU lg2 / sqrt / sin / cos) -> add.
// ----------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) add_noise_kernel(const T* __restrict__ x, T* __restrict__ out, long long groups,
                                                        float stddev, uint2 key) {
  pdl_launch_dependents();   // the next kernel in the stream (a PDL-launched GEMM) may start its prologue
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (long long)gridDim.x * blockDim.x) {
    float v[8];
    V8<T>::load(x + g * 8, v);
    const unsigned long long q = (unsigned long long)g * 2;
    const uint4 a = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)(q >> 32), 0x6e6f6973u, 0u), key);
    const uint4 b = philox4x32_10(make_uint4((uint32_t)(q + 1), (uint32_t)((q + 1) >> 32), 0x6e6f6973u, 0u), key);
    const uint32_t bits[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      // u1 in (0, 1], u2 in [0, 1):  r = sqrt(-2 ln u1), (n0, n1) = r (cos, sin)(2 pi u2)
      const float u1 = ((float)(bits[j] >> 8) + 1.0f) * (1.0f / 16777216.0f);
      const float u2 = (float)(bits[j + 1] >> 8) * (1.0f / 16777216.0f);
      const float r = stddev * sqrtf(-1.3862943611198906f * __log2f(u1));   // -2 ln u = -2 ln2 log2 u
      float sn, cs;
      __sincosf(6.283185307179586f * u2, &sn, &cs);
      v[j] = fmaf(r, cs, v[j]);
      v[j + 1] = fmaf(r, sn, v[j + 1]);
    }
    V8<T>::store(out + g * 8, v);
  }
}

// ----------------------------------------------------------------------------------------------
// patch extraction (einops Rearrange, simple_vit.py:127-129 ; Conv2d(k=s=P) im2col, vit.py:237-242)
// one thread per 8 output columns (16/32-byte store)
// ----------------------------------------------------------------------------------------------
template <typename TI, typename TO, bool VEC>
__global__ void im2col_kernel(const TI* __restrict__ img, int B, int C, int H, int W, int ph, int pw,
                              int order, TO* __restrict__ out, long long ld, int rows_out,
                              int row_off) {
  pdl_launch_dependents();   // the next kernel in the stream (a PDL-launched GEMM) may start its prologue
  const int nh = H / ph, nw = W / pw;
  const int kdim = C * ph * pw;
  const int groups = (int)(ld / 8);
  const long long total = (long long)B * nh * nw * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int gidx = (int)(i % groups);
    const long long prow = i / groups;
    const int pwi = (int)(prow % nw);
    const int phi = (int)((prow / nw) % nh);
    const int b = (int)(prow / ((long long)nw * nh));
    float v[8];
    if (VEC) {
      // (c p1 p2) order with pw % 8 == 0 and W % 8 == 0: the 8 inputs of this group are contiguous in the image
      const int k = gidx * 8;
      if (k < kdim) {
        const int p2 = k % pw, p1 = (k / pw) % ph, c = k / (pw * ph);
        V8<TI>::load(img + (((long long)b * C + c) * H + (phi * ph + p1)) * W + (pwi * pw + p2), v);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = gidx * 8 + j;
        float val = 0.f;
        if (k < kdim) {
          int c, p1, p2;
          if (order == NRV_PATCH_P1P2C) { c = k % C; p2 = (k / C) % pw; p1 = k / (C * pw); }
          else { p2 = k % pw; p1 = (k / pw) % ph; c = k / (pw * ph); }
          val = to_f32(img[(((long long)b * C + c) * H + (phi * ph + p1)) * W + (pwi * pw + p2)]);
        }
        v[j] = val;
      }
    }
    const long long orow = (long long)b * rows_out + (prow - (long long)b * nh * nw) + row_off;
    V8<TO>::store(out + orow * ld + gidx * 8, v);
  }
  // rows [0, row_off) of every image (the class-token slots) are zero, so that dW = dx^T * patches and the
  // forward GEMM can run over all rows_out rows
  const long long ztotal = (long long)B * row_off * groups;
  float z[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) z[j] = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ztotal;
       i += (long long)gridDim.x * blockDim.x) {
    const int gidx = (int)(i % groups);
    const long long r = i / groups;
    const long long orow = (r / row_off) * rows_out + (r % row_off);
    V8<TO>::store(out + orow * ld + gidx * 8, z);
  }
}

// posemb_sincos_2d (simple_vit.py:15-28), fp32 math as in the reference
__global__ void posemb_sincos_kernel(float* __restrict__ out, int h, int w, int dim, float temperature) {
  const int q = dim / 4;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= h * w * q) return;
  const int t = i / q, j = i - t * q;
  const int y = t / w, x = t - y * w;
  const float omega = 1.0f / powf(temperature, (float)j / (float)(q - 1));
  const float xa = (float)x * omega, ya = (float)y * omega;
  float* o = out + (long long)t * dim;
  o[j] = sinf(xa);
  o[q + j] = cosf(xa);
  o[2 * q + j] = sinf(ya);
  o[3 * q + j] = cosf(ya);
}

// x[b,0,:] = cls + pos[0]   (vit.py:341-342 cat(class_token) ; vit.py:174 + pos_embedding)
template <typename T>
__global__ void cls_token_kernel(const float* __restrict__ cls, const float* __restrict__ pos,
                                 T* __restrict__ x, int B, int tokens, int dim) {
  pdl_launch_dependents();   // the next kernel in the stream (a PDL-launched GEMM) may start its prologue
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * dim) return;
  const int b = i / dim, d = i - b * dim;
  x[((long long)b * tokens) * dim + d] = from_f32<T>(cls[d] + (pos ? pos[d] : 0.f));
}

// dpos[t,d] += sum_b dx[b,t,d] ; dcls[d] += sum_b dx[b,0,d]
// one thread = 8 columns of one token over a slice of the batch (blockIdx.y); slices meet in fp32 atomics
template <typename T>
__global__ void posemb_bwd_kernel(const T* __restrict__ dx, int B, int tokens, int dim,
                                  float* __restrict__ dpos, float* __restrict__ dcls) {
  pdl_launch_dependents();   // the next kernel in the stream (a PDL-launched GEMM) may start its prologue
  const int groups = dim / 8;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over tokens*groups
  if (i >= tokens * groups) return;
  const int t = i / groups, d = (i - t * groups) * 8;
  const int per = (B + gridDim.y - 1) / gridDim.y;
  const int b0 = blockIdx.y * per, b1 = min(B, b0 + per);
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
#pragma unroll 4
  for (int b = b0; b < b1; ++b) {
    float v[8];
    V8<T>::load(dx + ((long long)b * tokens + t) * dim + d, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] += v[j];
  }
  if (b1 <= b0) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (dpos) atomicAdd(dpos + (long long)t * dim + d + j, s[j]);
    if (dcls && t == 0) atomicAdd(dcls + d + j, s[j]);
  }
}

// ----------------------------------------------------------------------------------------------
// pooling (x.mean(dim=1), simple_vit.py:146 ; x[:, 0], vit.py:347)
// ----------------------------------------------------------------------------------------------
template <typename T>
__global__ void pool_fwd_kernel(const T* __restrict__ x, T* __restrict__ pooled, int B, int N,
                                int D, int pool) {
  pdl_launch_dependents();   // the next kernel in the stream (a PDL-launched GEMM) may start its prologue
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int b = i / D, d = i - b * D;
  const T* xb = x + (long long)b * N * D + d;
  float s;
  if (pool == NRV_POOL_CLS) {
    s = to_f32(xb[0]);
  } else {
    s = 0.f;
    for (int t = 0; t < N; ++t) s += to_f32(xb[(long long)t * D]);
    s /= (float)N;
  }
  pooled[(long long)b * D + d] = from_f32<T>(s);
}

template <typename T>
__global__ void pool_bwd_kernel(const T* __restrict__ dpooled, T* __restrict__ dx, int B, int N,
                                int D, int pool) {
  pdl_launch_dependents();   // the next kernel in the stream (a PDL-launched GEMM) may start its prologue
  const int groups = D / 8;
  const long long total = (long long)B * N * groups;
  const float inv = 1.f / (float)N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    const long long row = i / groups;
    const int t = (int)(row % N);
    const int b = (int)(row / N);
    float o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (pool == NRV_POOL_MEAN) {
      V8<T>::load(dpooled + (long long)b * D + g * 8, o);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] *= inv;
    } else if (t == 0) {
      V8<T>::load(dpooled + (long long)b * D + g * 8, o);
    }
    V8<T>::store(dx + row * D + g * 8, o);
  }
}

// ----------------------------------------------------------------------------------------------
// softmax cross-entropy with label smoothing, fwd + bwd in one pass; one warp per sample.
// (F.cross_entropy(preds, y, label_smoothing=eps), examples/baseline.py:70)
//   loss_i = -(1-eps) log p[y] - eps/C sum_c log p[c];  dz = (p - (1-eps) 1[y] - eps/C) * scale / B
// ----------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) softmax_ce_kernel(const float* __restrict__ logits,
                                                          long long ldl,
                                                          const long long* __restrict__ labels,
                                                          float eps, float* __restrict__ loss_mean,
                                                          T* __restrict__ dlogits, long long ldd,
                                                          float grad_scale, int B, int C) {
  pdl_launch_dependents();   // the next kernel in the stream (a PDL-launched GEMM) may start its prologue
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= B) return;
  const float* z = logits + (long long)row * ldl;
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, z[c]);
  mx = warp_max(mx);
  float se = 0.f, sz = 0.f;
  for (int c = lane; c < C; c += 32) { se += expf(z[c] - mx); sz += z[c]; }
  se = warp_sum(se); sz = warp_sum(sz);
  const float lse = mx + logf(se);
  const int y = (int)labels[row];
  if (lane == 0 && loss_mean) {
    const float logp_y = z[y] - lse;
    const float sum_logp = sz - (float)C * lse;
    const float li = -(1.f - eps) * logp_y - (eps / (float)C) * sum_logp;
    atomicAdd(loss_mean, li / (float)B);
  }
  if (dlogits) {
    const float sc = grad_scale / (float)B;
    T* d = dlogits + (long long)row * ldd;
    for (int c = lane; c < (int)ldd; c += 32) {
      float gval = 0.f;
      if (c < C) gval = (expf(z[c] - lse) - (c == y ? (1.f - eps) : 0.f) - eps / (float)C) * sc;
      d[c] = from_f32<T>(gval);
    }
  }
}

// ----------------------------------------------------------------------------------------------
// fused AdamW over a flat fp32 segment + bf16 shadow write (torch.optim.AdamW, CIFAR100.py:90-97)
// 30 B/param: read p,m,v,g (16) ; write p,m,v (12) + bf16 shadow (2)
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, float* __restrict__ m,
                                                     float* __restrict__ v, const float* __restrict__ g,
                                                     bf16* __restrict__ shadow, long long n, float lr,
                                                     float beta1, float beta2, float eps, float wd,
                                                     float bc1, float bc2_sqrt, float grad_scale,
                                                     const float* __restrict__ grad_scale_dev) {
  pdl_launch_dependents();   // the next kernel in the stream (a PDL-launched GEMM) may start its prologue
  const float gs = grad_scale * (grad_scale_dev ? *grad_scale_dev : 1.f);
  const float step_size = lr / bc1;
  const float decay = 1.f - lr * wd;
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float pa[4] = {pp.x, pp.y, pp.z, pp.w}, ma[4] = {mm.x, mm.y, mm.z, mm.w};
    float va[4] = {vv.x, vv.y, vv.z, vv.w};
    const float ga[4] = {gg.x * gs, gg.y * gs, gg.z * gs, gg.w * gs};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      pa[j] *= decay;
      ma[j] = beta1 * ma[j] + (1.f - beta1) * ga[j];
      va[j] = beta2 * va[j] + (1.f - beta2) * ga[j] * ga[j];
      const float denom = sqrtf(va[j]) / bc2_sqrt + eps;
      pa[j] -= step_size * (ma[j] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pa[0], pa[1], pa[2], pa[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(va[0], va[1], va[2], va[3]);
    if (shadow)
      reinterpret_cast<uint2*>(shadow)[i] = make_uint2(pack_bf16(pa[0], pa[1]), pack_bf16(pa[2], pa[3]));
  }
  // tail (n % 4)
  const long long t = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) {
    float pa = p[t] * decay, ga = g[t] * gs;
    const float ma = beta1 * m[t] + (1.f - beta1) * ga;
    const float va = beta2 * v[t] + (1.f - beta2) * ga * ga;
    pa -= step_size * (ma / (sqrtf(va) / bc2_sqrt + eps));
    p[t] = pa; m[t] = ma; v[t] = va;
    if (shadow) shadow[t] = __float2bfloat16(pa);
  }
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
  pdl_launch_dependents();   // the next kernel in the stream (a PDL-launched GEMM) may start its prologue
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 f = reinterpret_cast<const float4*>(src)[i];
    reinterpret_cast<uint2*>(dst)[i] = make_uint2(pack_bf16(f.x, f.y), pack_bf16(f.z, f.w));
  }
  const long long t = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) dst[t] = __float2bfloat16(src[t]);
}

// bf16 -> fp32 (gradient buckets come back from the bf16 all-reduce)
__global__ void cast_f32_kernel(const bf16* __restrict__ src, float* __restrict__ dst, long long n) {
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const uint2 u = reinterpret_cast<const uint2*>(src)[i];
    const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
    reinterpret_cast<float4*>(dst)[i] = make_float4(a.x, a.y, b.x, b.y);
  }
  const long long t = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) dst[t] = __bfloat162float(src[t]);
}

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n,
                                                     float* __restrict__ out) {
  __shared__ float red[8];
  float s = 0.f;
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 f = reinterpret_cast<const float4*>(g)[i];
    s += f.x * f.x + f.y * f.y + f.z * f.z + f.w * f.w;
  }
  const long long t = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) s += g[t] * g[t];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float r = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
    r = warp_sum(r);
    if (threadIdx.x == 0) atomicAdd(out, r);
  }
}

__global__ void clip_coef_kernel(const float* sumsq, float max_norm, float extra_scale, float* coef) {
  const float total = sqrtf(*sumsq) * extra_scale;  // norm of the (scaled) gradient
  *coef = fminf(1.f, max_norm / (total + 1e-6f));
}

static inline int grid_for(long long work_items, int threads, int sms, int per_sm = 8) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = (long long)sms * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace nrv

using namespace nrv;

#define NRV_ENTRY()                 \
  do {                              \
    int _rc = require_init();       \
    if (_rc) return _rc;            \
  } while (0)

#define NRV_DTYPE_OK(dt, who) \
  NRV_REQUIRE((dt) == NRV_BF16 || (dt) == NRV_F32, who ": dtype must be NRV_BF16 or NRV_F32 (got %d)", (int)(dt))

// run `stmt` with type alias T bound to the activation dtype
#define NRV_DISPATCH(dt, ...)                    \
  do {                                           \
    if ((dt) == NRV_BF16) { using T = bf16; __VA_ARGS__; } \
    else { using T = float; __VA_ARGS__; }       \
  } while (0)

static int colsum_splits(long long rows, int cols) {
  const int strips = (cols + 255) / 256;
  const long long sms = num_sms() > 0 ? num_sms() : 148;
  long long s = (sms * 4 + strips - 1) / strips;
  const long long maxs = (rows + 63) / 64;
  if (s > maxs) s = maxs;
  if (s < 1) s = 1;
  return (int)s;
}


namespace nrv {

int colsum_rows(const void* x, long long ldx, long long rows, int cols, int dtype, int period, int skip,
                float* out, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  NRV_ENTRY();
  NRV_DTYPE_OK(dtype, "nrv_colsum");
  NRV_REQUIRE(x && out && workspace, "nrv_colsum: null pointer");
  NRV_REQUIRE(cols % 8 == 0 && ldx % 8 == 0, "nrv_colsum: cols and ldx must be multiples of 8");
  NRV_REQUIRE(period >= 1 && skip >= 0 && skip < period + 1, "nrv_colsum: bad row filter");
  if (rows <= 0) return NRV_OK;
  const int splits = colsum_splits(rows, cols);
  NRV_REQUIRE(workspace_bytes >= (size_t)splits * cols * sizeof(float), "nrv_colsum: workspace too small");
  dim3 grid((cols + 255) / 256, splits);
  NRV_DISPATCH(dtype, colsum_kernel<T><<<grid, 256, 0, st>>>((const T*)x, ldx, rows, cols, period, skip, (float*)workspace));
  NRV_CUDA(cudaGetLastError());
  colreduce_finalize<<<(cols + 31) / 32, 256, 0, st>>>((const float*)workspace, splits, 1, cols, out, nullptr, nullptr);
  count_launch(2);
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

// out[cols] += column sums of x in one kernel (no workspace); `out` must be 16-byte aligned
int colsum_atomic(const void* x, long long ldx, long long rows, int cols, int dtype, float* out, cudaStream_t st) {
  NRV_ENTRY();
  NRV_DTYPE_OK(dtype, "nrv_colsum");
  NRV_REQUIRE(x && out, "nrv_colsum: null pointer");
  NRV_REQUIRE(cols % 8 == 0 && ldx % 8 == 0 && (reinterpret_cast<uintptr_t>(out) % 16) == 0,
              "nrv_colsum: cols and ldx must be multiples of 8, out 16-byte aligned");
  if (rows <= 0) return NRV_OK;
  const int strips = (cols + 255) / 256;
  int grid = 4 * (num_sms() > 0 ? num_sms() : 148);
  while (grid > strips && (long long)(grid * 8 / strips) * 16 > rows) grid /= 2;   // at least ~16 rows per warp
  if (grid * 8 < strips) grid = (strips + 7) / 8;
  NRV_DISPATCH(dtype, colsum_atomic_kernel<T><<<grid, 256, 0, st>>>((const T*)x, ldx, rows, cols, out));
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

int im2col_rows(const void* img, int img_dtype, int B, int C, int H, int W, int ph, int pw, int order,
                void* patches, int out_dtype, long long ld, int rows_out, int row_off, cudaStream_t st) {
  NRV_ENTRY();
  NRV_DTYPE_OK(img_dtype, "nrv_im2col(img)");
  NRV_DTYPE_OK(out_dtype, "nrv_im2col(out)");
  NRV_REQUIRE(img && patches, "nrv_im2col: null pointer");
  NRV_REQUIRE(ph > 0 && pw > 0 && H % ph == 0 && W % pw == 0, "Image dimensions must be divisible by the patch size.");
  NRV_REQUIRE(ld % 8 == 0 && ld >= (long long)C * ph * pw, "nrv_im2col: ld must be a multiple of 8 and >= C*ph*pw");
  NRV_REQUIRE(order == NRV_PATCH_P1P2C || order == NRV_PATCH_CP1P2, "nrv_im2col: bad order");
  const long long items = (long long)B * (H / ph) * (W / pw) * (ld / 8);
  if (items <= 0) return NRV_OK;
  const int grid = grid_for(items, 256, num_sms(), 16);
  const int kdim = C * ph * pw;
  const bool vec = order == NRV_PATCH_CP1P2 && pw % 8 == 0 && W % 8 == 0 && kdim % 8 == 0 &&
                   (reinterpret_cast<uintptr_t>(img) % 32) == 0;
#define NRV_IM2COL(TI, VEC) \
  NRV_DISPATCH(out_dtype, im2col_kernel<TI, T, VEC><<<grid, 256, 0, st>>>((const TI*)img, B, C, H, W, ph, pw, order, (T*)patches, ld, rows_out, row_off))
  if (img_dtype == NRV_F32) {
    if (vec) { NRV_IM2COL(float, true); } else { NRV_IM2COL(float, false); }
  } else {
    if (vec) { NRV_IM2COL(bf16, true); } else { NRV_IM2COL(bf16, false); }
  }
#undef NRV_IM2COL
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

}  // namespace nrv

extern "C" {

int nrv_layernorm_fwd(const void* x, const float* gamma, const float* beta, float eps, void* y,
                      float* mean, float* rstd, long long rows, int dim, int dtype, void* stream) {
  NRV_ENTRY();
  NRV_DTYPE_OK(dtype, "nrv_layernorm_fwd");
  NRV_REQUIRE(x && gamma && beta && y, "nrv_layernorm_fwd: null pointer");
  NRV_REQUIRE(dim % 8 == 0 && dim > 0 && dim <= 2048, "nrv_layernorm_fwd: dim must be a multiple of 8, <= 2048 (got %d)", dim);
  if (rows <= 0) return NRV_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int nch = (dim + 255) / 256;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if (dtype == NRV_BF16 && dim % 256 == 0 && nch >= 2 && nch <= 4 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)y % 16) == 0) {
    // packed-math kernel, persistent: 8 CTAs of 8 warps per SM
    const unsigned pg = (unsigned)(grid < (unsigned)(num_sms() * 8) ? grid : (unsigned)(num_sms() * 8));
    if (nch == 2) ln_fwd_bf16_kernel<2><<<pg, 256, 0, st>>>((const bf16*)x, gamma, beta, eps, (bf16*)y, mean, rstd, rows);
    else if (nch == 3) ln_fwd_bf16_kernel<3><<<pg, 256, 0, st>>>((const bf16*)x, gamma, beta, eps, (bf16*)y, mean, rstd, rows);
    else ln_fwd_bf16_kernel<4><<<pg, 256, 0, st>>>((const bf16*)x, gamma, beta, eps, (bf16*)y, mean, rstd, rows);
    count_launch();
    NRV_CUDA(cudaGetLastError());
    return NRV_OK;
  }
#define LAUNCH_LNF(N) ln_fwd_kernel<T, N><<<grid, 256, 0, st>>>((const T*)x, gamma, beta, eps, (T*)y, mean, rstd, rows, dim)
  NRV_DISPATCH(dtype, switch (nch) {
    case 1: LAUNCH_LNF(1); break; case 2: LAUNCH_LNF(2); break; case 3: LAUNCH_LNF(3); break;
    case 4: LAUNCH_LNF(4); break; case 5: LAUNCH_LNF(5); break; case 6: LAUNCH_LNF(6); break;
    case 7: LAUNCH_LNF(7); break; default: LAUNCH_LNF(8); break;
  });
#undef LAUNCH_LNF
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

static int ln_bwd_blocks(long long rows, int dim) {
  const int nw = (dim + 255) / 256;
  const int rpb = LNB_WARPS / nw;
  long long b = (rows + rpb - 1) / rpb;
  const long long sms = num_sms() > 0 ? num_sms() : 148;
  const long long cap = sms * 2;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

size_t nrv_layernorm_bwd_workspace(long long rows, int dim) {
  const long long sms = num_sms() > 0 ? num_sms() : 148;
  (void)rows;
  return (size_t)(sms * 2) * 3 * dim * sizeof(float);
}

int nrv_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd,
                      const float* gamma, const void* dres, void* dx, float* dgamma, float* dbeta,
                      float* colsum, const float* beta, void* xn_out, long long rows, int dim, int dtype, void* workspace,
                      size_t workspace_bytes, void* stream) {
  NRV_ENTRY();
  NRV_DTYPE_OK(dtype, "nrv_layernorm_bwd");
  NRV_REQUIRE(dy && x && mean && rstd && gamma && dx && workspace, "nrv_layernorm_bwd: null pointer");
  NRV_REQUIRE(xn_out == nullptr || beta != nullptr, "nrv_layernorm_bwd: xn_out needs beta");
  NRV_REQUIRE(dim % 8 == 0 && dim > 0 && dim <= 1536, "nrv_layernorm_bwd: dim must be a multiple of 8, <= 1536 (got %d)", dim);
  if (rows <= 0) return NRV_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = ln_bwd_blocks(rows, dim);
  NRV_REQUIRE(workspace_bytes >= (size_t)blocks * 3 * dim * sizeof(float), "nrv_layernorm_bwd: workspace too small");
  const int nw = (dim + 255) / 256;
  const size_t smem = 3 * (size_t)dim * sizeof(float);
  // bf16, dim = 512 / 768 / 1024, 16-byte aligned operands: the bulk-copy staged kernel
  const bool tma_ok = dtype == NRV_BF16 && dim % 256 == 0 && nw >= 2 && nw <= 4 && rows >= 64 &&
                      ((uintptr_t)dy % 16) == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)dx % 16) == 0 &&
                      (dres == nullptr || ((uintptr_t)dres % 16) == 0) && ((uintptr_t)dgamma % 16) == 0 &&
                      ((uintptr_t)dbeta % 16) == 0 && ((uintptr_t)colsum % 16) == 0 && ((uintptr_t)mean % 16) == 0 &&
                      ((uintptr_t)rstd % 16) == 0 && ((uintptr_t)xn_out % 16) == 0;
  if (tma_ok) {
    const long long chunks = (rows + LNT_ROWS - 1) / LNT_ROWS;
    const int tb = (int)(chunks < (long long)blocks ? chunks : (long long)blocks);
#define NRV_LNT(CPL, XN) launch_ln_bwd_tma<CPL, XN>(dy, x, mean, rstd, gamma, dres, dx, dgamma, dbeta, colsum, rows, tb, st, beta, xn_out)
    int rc;
    if (xn_out != nullptr) rc = nw == 2 ? NRV_LNT(2, true) : nw == 3 ? NRV_LNT(3, true) : NRV_LNT(4, true);
    else rc = nw == 2 ? NRV_LNT(2, false) : nw == 3 ? NRV_LNT(3, false) : NRV_LNT(4, false);
#undef NRV_LNT
    if (rc) return rc;
    count_launch();
    NRV_CUDA(cudaGetLastError());
    return NRV_OK;
  }
#define LAUNCH_LNB(N) ln_bwd_kernel<T, N><<<blocks, LNB_WARPS * 32, smem, st>>>((const T*)dy, (const T*)x, mean, rstd, gamma, (const T*)dres, (T*)dx, (float*)workspace, rows, dim, beta, (T*)xn_out)
  NRV_DISPATCH(dtype, switch (nw) {
    case 1: LAUNCH_LNB(1); break; case 2: LAUNCH_LNB(2); break; case 3: LAUNCH_LNB(3); break;
    case 4: LAUNCH_LNB(4); break; case 5: LAUNCH_LNB(5); break; default: LAUNCH_LNB(6); break;
  });
#undef LAUNCH_LNB
  NRV_CUDA(cudaGetLastError());
  const int total = 3 * dim;
  colreduce_finalize<<<(total + 31) / 32, 256, 0, st>>>((const float*)workspace, blocks, 3, dim, dgamma, dbeta, colsum);
  count_launch(2);
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

size_t nrv_colsum_workspace(long long rows, int cols) {
  return (size_t)colsum_splits(rows, cols) * cols * sizeof(float);
}

int nrv_colsum(const void* x, long long ldx, long long rows, int cols, int dtype, float* out,
               void* workspace, size_t workspace_bytes, void* stream) {
  return colsum_rows(x, ldx, rows, cols, dtype, 1, 0, out, workspace, workspace_bytes, (cudaStream_t)stream);
}

int nrv_im2col(const void* img, int img_dtype, int B, int C, int H, int W, int ph, int pw,
               int order, void* patches, int out_dtype, long long ld, void* stream) {
  const int n = (ph > 0 && pw > 0) ? (H / ph) * (W / pw) : 0;
  return im2col_rows(img, img_dtype, B, C, H, W, ph, pw, order, patches, out_dtype, ld, n, 0, (cudaStream_t)stream);
}

int nrv_dropout(const void* x, const void* residual, void* out, long long n, int dtype, float p,
                unsigned long long seed, int layer, int site, void* stream) {
  NRV_ENTRY();
  NRV_DTYPE_OK(dtype, "nrv_dropout");
  NRV_REQUIRE(x && out, "nrv_dropout: null pointer");
  NRV_REQUIRE(n % 8 == 0, "nrv_dropout: n must be a multiple of 8 (got %lld)", n);
  NRV_REQUIRE(p >= 0.f && p < 1.f, "nrv_dropout: p must be in [0, 1) (got %g)", (double)p);
  NRV_REQUIRE(layer >= -1 && site >= 0 && site < 8, "nrv_dropout: bad (layer, site)");
  if (n <= 0) return NRV_OK;
  const long long groups = n / 8;
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  const uint32_t stream_id = (uint32_t)(layer + 1) * 8u + (uint32_t)site;
  const int grid = grid_for(groups, 256, num_sms(), 8);
  NRV_DISPATCH(dtype, dropout_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)x, (const T*)residual, (T*)out, groups, p,
                                                                                  1.f / (1.f - p), key, stream_id));
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

int nrv_add_gaussian_noise(const void* x, void* out, long long n, int dtype, float stddev, unsigned long long seed,
                           void* stream) {
  NRV_ENTRY();
  NRV_DTYPE_OK(dtype, "nrv_add_gaussian_noise");
  NRV_REQUIRE(x && out, "nrv_add_gaussian_noise: null pointer");
  NRV_REQUIRE(n % 8 == 0, "nrv_add_gaussian_noise: n must be a multiple of 8 (got %lld)", n);
  if (n <= 0) return NRV_OK;
  const long long groups = n / 8;
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  NRV_DISPATCH(dtype, add_noise_kernel<T><<<grid_for(groups, 256, num_sms(), 8), 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)out, groups,
                                                                                                              stddev, key));
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

int nrv_posemb_sincos_2d(float* out, int h, int w, int dim, float temperature, void* stream) {
  NRV_ENTRY();
  NRV_REQUIRE(out != nullptr && h > 0 && w > 0, "nrv_posemb_sincos_2d: bad arguments");
  NRV_REQUIRE(dim % 4 == 0, "feature dimension must be multiple of 4 for sincos emb");
  NRV_REQUIRE(dim > 4, "nrv_posemb_sincos_2d: dim must be > 4 (dim/4 - 1 divides)");
  const int total = h * w * (dim / 4);
  posemb_sincos_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(out, h, w, dim, temperature);
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

int nrv_cls_token_fwd(const float* cls, const float* pos, void* x, int B, int tokens, int dim,
                      int dtype, void* stream) {
  NRV_ENTRY();
  NRV_DTYPE_OK(dtype, "nrv_cls_token_fwd");
  NRV_REQUIRE(cls && x, "nrv_cls_token_fwd: null pointer");
  if (B <= 0) return NRV_OK;
  NRV_DISPATCH(dtype, cls_token_kernel<T><<<(B * dim + 255) / 256, 256, 0, (cudaStream_t)stream>>>(cls, pos, (T*)x, B, tokens, dim));
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

int nrv_posemb_bwd(const void* dx, int B, int tokens, int dim, int dtype, float* dpos, float* dcls,
                   void* stream) {
  NRV_ENTRY();
  NRV_DTYPE_OK(dtype, "nrv_posemb_bwd");
  NRV_REQUIRE(dx, "nrv_posemb_bwd: null pointer");
  if (B <= 0 || (!dpos && !dcls)) return NRV_OK;
  NRV_REQUIRE(dim % 8 == 0, "nrv_posemb_bwd: dim must be a multiple of 8");
  const int work = (dpos ? tokens : 1) * (dim / 8);
  const dim3 grid((work + 127) / 128, B >= 64 ? 8 : 1);
  NRV_DISPATCH(dtype, posemb_bwd_kernel<T><<<grid, 128, 0, (cudaStream_t)stream>>>((const T*)dx, B, tokens, dim, dpos, dcls));
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

int nrv_pool_fwd(const void* x, void* pooled, int B, int N, int D, int pool, int dtype, void* stream) {
  NRV_ENTRY();
  NRV_DTYPE_OK(dtype, "nrv_pool_fwd");
  NRV_REQUIRE(x && pooled, "nrv_pool_fwd: null pointer");
  NRV_REQUIRE(pool == NRV_POOL_MEAN || pool == NRV_POOL_CLS, "nrv_pool_fwd: bad pool mode");
  if (B <= 0) return NRV_OK;
  NRV_DISPATCH(dtype, pool_fwd_kernel<T><<<(B * D + 127) / 128, 128, 0, (cudaStream_t)stream>>>((const T*)x, (T*)pooled, B, N, D, pool));
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

int nrv_pool_bwd(const void* dpooled, void* dx, int B, int N, int D, int pool, int dtype, void* stream) {
  NRV_ENTRY();
  NRV_DTYPE_OK(dtype, "nrv_pool_bwd");
  NRV_REQUIRE(dpooled && dx, "nrv_pool_bwd: null pointer");
  NRV_REQUIRE(D % 8 == 0, "nrv_pool_bwd: D must be a multiple of 8");
  NRV_REQUIRE(pool == NRV_POOL_MEAN || pool == NRV_POOL_CLS, "nrv_pool_bwd: bad pool mode");
  if (B <= 0) return NRV_OK;
  const long long items = (long long)B * N * (D / 8);
  NRV_DISPATCH(dtype, pool_bwd_kernel<T><<<grid_for(items, 256, num_sms(), 16), 256, 0, (cudaStream_t)stream>>>((const T*)dpooled, (T*)dx, B, N, D, pool));
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

int nrv_softmax_ce(const float* logits, long long ldl, const long long* labels, float label_smoothing,
                   float* loss_mean, void* dlogits, int dl_dtype, long long ldd, float grad_scale,
                   int B, int C, void* stream) {
  NRV_ENTRY();
  NRV_DTYPE_OK(dl_dtype, "nrv_softmax_ce");
  NRV_REQUIRE(logits && labels, "nrv_softmax_ce: null pointer");
  NRV_REQUIRE(ldl >= C && (!dlogits || ldd >= C), "nrv_softmax_ce: leading dims must be >= C");
  if (B <= 0) return NRV_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (loss_mean) NRV_CUDA(cudaMemsetAsync(loss_mean, 0, sizeof(float), st));
  NRV_DISPATCH(dl_dtype, softmax_ce_kernel<T><<<(B + 7) / 8, 256, 0, st>>>(logits, ldl, labels, label_smoothing, loss_mean, (T*)dlogits, ldd, grad_scale, B, C));
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

int nrv_adamw(float* p, float* m, float* v, const float* g, void* shadow, long long n, float lr,
              float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
              const float* grad_scale_dev, void* stream) {
  NRV_ENTRY();
  NRV_REQUIRE(p && m && v && g, "nrv_adamw: null pointer");
  NRV_REQUIRE(step >= 1, "nrv_adamw: step must be >= 1");
  NRV_REQUIRE(((uintptr_t)p % 16) == 0 && ((uintptr_t)m % 16) == 0 && ((uintptr_t)v % 16) == 0 && ((uintptr_t)g % 16) == 0 && ((uintptr_t)shadow % 8) == 0,
              "nrv_adamw: buffers must be 16-byte aligned (shadow 8-byte)");
  if (n <= 0) return NRV_OK;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const int grid = grid_for((n + 3) / 4, 256, num_sms(), 8);
  adamw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, m, v, g, (bf16*)shadow, n, lr, beta1, beta2, eps, weight_decay, (float)bc1, (float)sqrt(bc2), grad_scale, grad_scale_dev);
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

int nrv_cast_bf16(const float* src, void* dst, long long n, void* stream) {
  NRV_ENTRY();
  NRV_REQUIRE(src && dst, "nrv_cast_bf16: null pointer");
  NRV_REQUIRE(((uintptr_t)src % 16) == 0 && ((uintptr_t)dst % 8) == 0, "nrv_cast_bf16: alignment");
  if (n <= 0) return NRV_OK;
  cast_bf16_kernel<<<grid_for((n + 3) / 4, 256, num_sms(), 8), 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, n);
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

int nrv_cast_f32(const void* src, float* dst, long long n, void* stream) {
  NRV_ENTRY();
  NRV_REQUIRE(src && dst, "nrv_cast_f32: null pointer");
  NRV_REQUIRE(((uintptr_t)dst % 16) == 0 && ((uintptr_t)src % 8) == 0, "nrv_cast_f32: alignment");
  if (n <= 0) return NRV_OK;
  cast_f32_kernel<<<grid_for((n + 3) / 4, 256, num_sms(), 8), 256, 0, (cudaStream_t)stream>>>((const bf16*)src, dst, n);
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

int nrv_sumsq(const float* g, long long n, float* out, void* stream) {
  NRV_ENTRY();
  NRV_REQUIRE(g && out, "nrv_sumsq: null pointer");
  NRV_REQUIRE(((uintptr_t)g % 16) == 0, "nrv_sumsq: alignment");
  if (n <= 0) return NRV_OK;
  sumsq_kernel<<<grid_for((n + 3) / 4, 256, num_sms(), 4), 256, 0, (cudaStream_t)stream>>>(g, n, out);
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

int nrv_clip_coef(const float* sumsq, float max_norm, float extra_scale, float* coef, void* stream) {
  NRV_ENTRY();
  NRV_REQUIRE(sumsq && coef, "nrv_clip_coef: null pointer");
  clip_coef_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(sumsq, max_norm, extra_scale, coef);
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

}  // extern "C"
