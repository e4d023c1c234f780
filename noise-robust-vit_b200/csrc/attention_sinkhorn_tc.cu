// robust=True attention on the tensor cores (tcgen05), forward: softmax followed by the reference's 3-iteration
// Sinkhorn normalisation (utils.py:1025-1037, selected at simple_vit.py:56-57 ; vit.py:98-110), bf16, dh = 64,
// up to 208 tokens (ViT-B/16 and ViT-L/16: 197; the CIFAR / README SimpleViT configs: 64).  Other shapes and the fp32 check
// mode run the CUDA-core kernels of attention_sinkhorn.cu, which also define the statistics layout shared with backward.
//
// Every normalisation step multiplies rows or columns by a scalar, so the final matrix is
//     P = diag(a) E diag(b),      E_ij = exp(s_ij - max_i)          (fixed; bf16, <= 1)
// and the seven steps only update the two vectors:   row step  a_i = 1 / (E b)_i ,   column step  b_j = 1 / (E^T a)_j .
// The kernel therefore
//   1. S = Q K^T on the tensor cores (two 128-row tiles of one (batch, head) item into TMEM),
//   2. thread = score row: max, exp2, E written ONCE as bf16 into shared memory in the K-major SWIZZLE_128B operand
//      layout (rows and columns beyond N are zero),
//   3. runs the 7 matrix-vector passes on E from shared memory (row passes: thread = row; column passes: thread = column,
//      both conflict-free in the swizzled layout), recording the row / column sums of every step exactly as the
//      reference's step-by-step loop would see them ([B,H,8,N] statistics: lse + 7 sum vectors),
//   4. scales V's rows by b in place and computes O' = E (b o V) on the tensor cores (A and B from shared memory),
//   5. scales the rows of O' by a in the epilogue.
// The matrix is never renormalised in place, so no rounding accumulates over the steps: the only quantisation is E to
// bf16, as in the softmax kernels.  One CTA per SM, 8 arithmetic warps + 1 control warp (TMA + MMA issue).
#include "common.cuh"
#include "nrvit_internal.h"

namespace nrv {

constexpr int SKT_THREADS = 32 * 9;
constexpr int SKT_ROWS = 256;                  // two M = 128 tiles
constexpr int SKT_CHUNK = SKT_ROWS * 128;      // bytes of one 64-key chunk of E
constexpr int SKT_MAXN = 208;

struct SkTcParams {
  int B, N, H, NP, items;
  float scale, scale_log2e;
  bf16* out;       // [B, N, H*64]
  float* stats;    // [B, H, 8, N]
};

// named barrier over the 8 arithmetic warps
__device__ __forceinline__ void skt_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// y_j = sum_i M_ij x_i over the bf16 matrix in the K-major SWIZZLE_128B layout (chunks of 64 columns, `chunk` bytes
// apart; row pitch 128 B).  Threads 0..207 of the 8 arithmetic warps: thread = (column pair, row half) -- a 32-bit load
// fetches two columns, the 32 lanes of a warp cover one whole 128-byte row of a chunk (conflict-free under the swizzle),
// and the two halves meet through `partial` ([2][256] floats).  Call from all 256 threads; returns the sum for
// column `r` (valid for r < NP) after one named-barrier round.
__device__ __forceinline__ float skt_colsum(uint32_t sM, uint32_t chunk, int N, int NP, uint32_t x_addr, float* partial, int tid, int r) {
  if (tid < 208) {
    const int cp = tid % 104, half = tid / 104;
    const int col = 2 * cp;
    if (col < NP) {
      const int nh = ((N + 1) / 2 + 7) & ~7;                 // rows of the first half, a multiple of 8
      const int i0 = half * nh, i1 = half == 0 ? (nh < N ? nh : N) : N;
      const uint32_t base = sM + (uint32_t)(col >> 6) * chunk + (col & 7) * 2;
      const int u = (col & 63) >> 3;
      // i0 is a multiple of 8, so row i + k of a group of eight sits in swizzle phase k: the eight in-atom offsets are
      // computed once (the ncu source view had the per-load address arithmetic at 4 of the loop's ~10 instructions per row)
      uint32_t off[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) off[k] = base + k * 128 + ((u ^ k) << 4);
      float a0 = 0.f, a1 = 0.f, c0 = 0.f, c1 = 0.f;        // two independent chains per column
      int i = i0;
      for (; i + 8 <= i1; i += 8) {
        const uint32_t rb = (uint32_t)i * 128;
        const float4 xa = lds128f(x_addr + i * 4), xb = lds128f(x_addr + i * 4 + 16);
        const float xs[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
        uint32_t w[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w[k]) : "r"(off[k] + rb) : "memory");
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
          a0 = fmaf(__uint_as_float(w[k] << 16), xs[k], a0);
          a1 = fmaf(__uint_as_float(w[k] & 0xffff0000u), xs[k], a1);
          c0 = fmaf(__uint_as_float(w[k + 1] << 16), xs[k + 1], c0);
          c1 = fmaf(__uint_as_float(w[k + 1] & 0xffff0000u), xs[k + 1], c1);
        }
      }
      for (; i < i1; ++i) {
        uint32_t w;
        float xk;
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(base + i * 128 + ((u ^ (i & 7)) << 4)) : "memory");
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(xk) : "r"(x_addr + i * 4) : "memory");
        a0 = fmaf(__uint_as_float(w << 16), xk, a0);
        a1 = fmaf(__uint_as_float(w & 0xffff0000u), xk, a1);
      }
      a0 += c0; a1 += c1;
      *reinterpret_cast<float2*>(partial + half * 256 + col) = make_float2(a0, a1);
    }
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
  return r < NP ? partial[r] + partial[256 + r] : 0.f;
}

__global__ void __launch_bounds__(SKT_THREADS, 1)
sinkhorn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tm128, const __grid_constant__ CUtensorMap tm16, const SkTcParams p) {
  pdl_launch_dependents();   // the next kernel in the stream (a PDL-launched GEMM) may start its prologue
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const int NP = p.NP, N = p.N, H = p.H;
  const int KVB = NP * 128;                                   // bytes of the K / V tile
  const int KVA = (KVB + 1023) & ~1023;
  const uint32_t sE = sbase;                                  // [4 chunks][256 rows][128 B]
  const uint32_t sQ = sE + 4 * SKT_CHUNK;                     // [256 rows][128 B]
  const uint32_t sK = sQ + SKT_ROWS * 128;
  const uint32_t sV = sK + KVA;
  const uint32_t sVec = sV + KVA;                             // a[256] | b[256] | column-sum partials [2][256], fp32
  const uint32_t bar0 = sVec + 4 * SKT_ROWS * 4;
  const uint32_t bar_ld = bar0, bar_s = bar0 + 8, bar_o = bar0 + 16, bar_v = bar0 + 24, bar_free = bar0 + 32;
  float* vec_a = reinterpret_cast<float*>(smem + (sVec - sbase));
  float* vec_b = vec_a + SKT_ROWS;
  float* vec_p = vec_b + SKT_ROWS;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + (bar0 - sbase) + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 8) {
    if (elect_one()) {
      tma_prefetch_desc(&tm128);
      tma_prefetch_desc(&tm16);
      mbar_init(bar_ld, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_o, 1);
      mbar_init(bar_v, 8);        // the 8 arithmetic warps: V rows scaled, E complete
      mbar_init(bar_free, 8);     // the 8 arithmetic warps: O' read out of tensor memory
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t T = *tmem_ptr_smem;
  const int my_items = ((int)blockIdx.x < p.items) ? (p.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int ksteps_n = NP / 16;
  const int ntiles = (N + 127) / 128;     // row tiles that hold queries (the second one is skipped up to 128 tokens)

  if (warp == 8) {
    // ================================ TMA + MMA issue (one lane) ================================
    if (elect_one()) {
      const uint64_t dfix = make_smem_desc_sw128(0, 16, 1024);
      const uint64_t dfix_v = make_smem_desc_sw128(0, (uint32_t)KVA, 1024);     // MN-major B: rows = keys, 64 columns of d
      const uint32_t idesc_s = make_idesc(1u, 0u, 0u, 128u, (uint32_t)NP);
      const uint32_t idesc_o = make_idesc(1u, 0u, 1u, 128u, 64u);
      auto load_rows = [&](uint32_t dst, int which, int h, int row0, int nrows, int b) {
        int r = 0;
        for (; r + 128 <= nrows; r += 128) tma_load_4d(dst + r * 128, &tm128, bar_ld, 0, which * H + h, row0 + r, b);
        for (; r < nrows; r += 16) tma_load_4d(dst + r * 128, &tm16, bar_ld, 0, which * H + h, row0 + r, b);
      };
      for (int li = 0; li < my_items; ++li) {
        const int item = (int)blockIdx.x + li * (int)gridDim.x;
        const int b = item / H, h = item % H;
        const uint32_t ph = li & 1;
        // Q (both row tiles; rows beyond N are out of bounds and arrive as zeros), K, V.  The previous item's MMAs have
        // retired (bar_o was waited) and its epilogue does not read shared memory.
        mbar_arrive_expect_tx(bar_ld, ntiles * 16384 + 2 * KVB);
        load_rows(sQ, 0, h, 0, ntiles * 128, b);
        load_rows(sK, 1, h, 0, NP, b);
        load_rows(sV, 2, h, 0, NP, b);
        mbar_wait(bar_ld, ph, 40);
        if (li > 0) mbar_wait(bar_free, (li - 1) & 1, 43);   // the previous item's O' has left tensor memory
        tc_fence_after();
        for (int t = 0; t < ntiles; ++t)
          for (int ks = 0; ks < 4; ++ks)
            umma_bf16(T + t * 256, dfix + (uint64_t)((sQ + t * 16384 + ks * 32) >> 4), dfix + (uint64_t)((sK + ks * 32) >> 4),
                      idesc_s, ks > 0);
        umma_commit(bar_s);
        // E in shared memory, V scaled by b: O' = E (b o V), both row tiles (the score columns are free again)
        mbar_wait(bar_v, ph, 41);
        tc_fence_after();
        for (int t = 0; t < ntiles; ++t)
          for (int ks = 0; ks < ksteps_n; ++ks)
            umma_bf16(T + t * 256, dfix + (uint64_t)((sE + (ks >> 2) * SKT_CHUNK + t * 16384 + (ks & 3) * 32) >> 4),
                      dfix_v + (uint64_t)((sV + ks * 2048) >> 4), idesc_o, ks > 0);
        umma_commit(bar_o);
        mbar_wait(bar_o, ph, 42);
      }
    }
    __syncwarp();
  } else {
    // ================================ arithmetic warps: thread = row r ==========================
    const int tile = warp >> 2, q = warp & 3;
    const int r = tile * 128 + q * 32 + lane;              // query row (and key column in the column passes)
    const int tid = threadIdx.x;                           // 0..255
    const uint32_t T_S = T + tile * 256 + ((uint32_t)(q * 32) << 16);
    const int nch = NP / 16;
    const uint64_t c2 = f2_pack(p.scale_log2e, p.scale_log2e);
    const uint32_t e_row = sE + r * 128;
    for (int li = 0; li < my_items; ++li) {
      const int item = (int)blockIdx.x + li * (int)gridDim.x;
      const int b = item / H, h = item % H;
      const uint32_t ph = li & 1;
      float* st = p.stats + (long long)item * 8 * N;
      mbar_wait(bar_s, ph, 50);
      tc_fence_after();
      // ---- row max, E = exp2(s c - max c) as bf16 into the operand layout, a0 = 1 / sum
      const bool row_ok = r < N;
      float mx = 0.f, tot = 1.f;
      {
        float m0 = -INFINITY, m1 = -INFINITY;
        for (int c0 = 0; c0 < nch; c0 += 4) {
          uint32_t v[4][16];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (c0 + k < nch) tmem_ld_32x16(T_S + (c0 + k) * 16, v[k]);
          tmem_wait_ld();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (c0 + k < nch) f2_max16(v[k], (c0 + k) * 16, N, m0, m1);
        }
        mx = fmaxf(m0, m1);
        const float noff = -mx * p.scale_log2e;
        const uint64_t noff2 = f2_pack(noff, noff);
        uint64_t sum2 = f2_pack(0.f, 0.f);
        for (int c0 = 0; c0 < 16; c0 += 4) {           // all 16 chunks of 16 keys (256): columns beyond NP are written as zeros
          uint32_t v[4][16];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (c0 + k < nch) tmem_ld_32x16(T_S + (c0 + k) * 16, v[k]);
          tmem_wait_ld();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int ch = c0 + k, j0 = ch * 16;
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              float e0 = 0.f, e1 = 0.f;
              if (ch < nch && row_ok) {
                float x0, x1;
                f2_unpack(f2_fma(f2_pack(__uint_as_float(v[k][j]), __uint_as_float(v[k][j + 1])), c2, noff2), x0, x1);
                e0 = j0 + j < N ? ex2f(x0) : 0.f;
                e1 = j0 + j + 1 < N ? ex2f(x1) : 0.f;
              }
              pk[j >> 1] = pack_bf16(e0, e1);
              // the sums below run over E as the tensor cores and the passes will read it
              const float2 rr = unpack_bf16(pk[j >> 1]);
              sum2 = f2_add(sum2, f2_pack(rr.x, rr.y));
            }
            // keys j0 .. j0+15 = 16-byte units (j0 % 64) / 8 and +1 of chunk j0 / 64
            const uint32_t base = e_row + (uint32_t)(j0 >> 6) * SKT_CHUNK;
            const int u = (j0 & 63) >> 3;
            sts128(base + (((u) ^ (r & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
            sts128(base + (((u + 1) ^ (r & 7)) << 4), pk[4], pk[5], pk[6], pk[7]);
          }
        }
        float s0, s1;
        f2_unpack(sum2, s0, s1);
        tot = s0 + s1;
      }
      if (row_ok) st[r] = mx * p.scale + __logf(tot);         // lse: P0 = exp(s - lse)
      vec_a[r] = row_ok ? 1.f / tot : 0.f;                     // P0 = a0 o E
      vec_b[r] = 1.f;
      tc_fence_before();
      skt_sync();
      // ---- the 7 normalisation steps on the two scaling vectors
      for (int k = 0; k < 7; ++k) {
        if ((k & 1) == 0) {
          // row step: s_i = a_i (E b)_i ; a_i <- 1 / (E b)_i
          float acc = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
          for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int j0 = c * 64 + u * 8;
              if (j0 < NP) {
                const uint4 w = lds128(e_row + c * SKT_CHUNK + ((u ^ (r & 7)) << 4));
                const float4 b0 = lds128f(smem_u32(vec_b + j0)), b1 = lds128f(smem_u32(vec_b + j0 + 4));
                const float2 e0 = unpack_bf16(w.x), e1 = unpack_bf16(w.y), e2 = unpack_bf16(w.z), e3 = unpack_bf16(w.w);
                acc = fmaf(e0.x, b0.x, acc); acc1 = fmaf(e0.y, b0.y, acc1); acc2 = fmaf(e1.x, b0.z, acc2); acc3 = fmaf(e1.y, b0.w, acc3);
                acc = fmaf(e2.x, b1.x, acc); acc1 = fmaf(e2.y, b1.y, acc1); acc2 = fmaf(e3.x, b1.z, acc2); acc3 = fmaf(e3.y, b1.w, acc3);
              }
            }
          }
          acc = (acc + acc1) + (acc2 + acc3);                 // (four chains instead of one 8-deep dependent FMA chain per unit)
          if (row_ok) {
            st[(1 + k) * N + r] = vec_a[r] * acc;
            vec_a[r] = 1.f / acc;
          }
        } else {
          // column step: s_j = b_j (E^T a)_j ; b_j <- 1 / (E^T a)_j
          const float acc = skt_colsum(sE, SKT_CHUNK, N, NP, smem_u32(vec_a), vec_p, tid, r);
          if (r < N) {
            st[(1 + k) * N + r] = vec_b[r] * acc;
            vec_b[r] = 1.f / acc;
          } else {
            vec_b[r] = 0.f;       // padding keys: E is zero there
          }
        }
        skt_sync();
      }
      // ---- V rows scaled by b in place (bf16), then O' = E (b o V) on the tensor cores
      if (r < NP) {
        const float bj = vec_b[r];
        const uint64_t b2 = f2_pack(bj, bj);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint32_t addr = sV + r * 128 + (u << 4);
          const uint4 w = lds128(addr);
          uint32_t o[4];
          const uint32_t in[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = unpack_bf16(in[e]);
            float x0, x1;
            f2_unpack(f2_mul(f2_pack(f.x, f.y), b2), x0, x1);
            o[e] = pack_bf16(x0, x1);
          }
          sts128(addr, o[0], o[1], o[2], o[3]);
        }
      }
      const float ai = vec_a[r];
      fence_async_smem();          // E and b o V were written by the generic proxy: visible to the tensor cores
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_v);
      mbar_wait(bar_o, ph, 51);
      tc_fence_after();
      // ---- epilogue: out row = a_i * O'_i
      {
        bf16* dst = p.out + ((long long)b * N + r) * (H * 64) + (long long)h * 64;
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t v[2][16];
          tmem_ld_32x16(T_S + c0, v[0]);
          tmem_ld_32x16(T_S + c0 + 16, v[1]);
          tmem_wait_ld();
          if (row_ok) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              uint32_t w[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) w[e] = pack_bf16(__uint_as_float(v[k][2 * e]) * ai, __uint_as_float(v[k][2 * e + 1]) * ai);
              *reinterpret_cast<uint4*>(dst + c0 + 16 * k) = make_uint4(w[0], w[1], w[2], w[3]);
              *reinterpret_cast<uint4*>(dst + c0 + 16 * k + 8) = make_uint4(w[4], w[5], w[6], w[7]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_free);   // this warp's O' rows are in registers: the next item's S may overwrite them
      skt_sync();                             // a / b are rewritten at the start of the next item
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(T, 512);
}

// ================================================================================================================
// Backward.  With E = P0 = softmax(S) (recomputed from q, k and the stashed lse) the forward pass is, in terms of the
// stashed step sums s_0..s_6 (attention_sinkhorn.cu):
//   a1 = 1/s0, b1 = 1/s1, a2 = a1/s2, b2 = b1/s3, a3 = a2/s4, b3 = b2/s5, a4 = a3/s6,     P = a4 o E o b3,
// where every row step is a = 1 / (E b) and every column step b = 1 / (E^T a).  Differentiating that chain gives, with
// G = dO V^T, delta_i = <dO_i, O_i>, M1 = P'^T dO for P' = a4 o E:
//   dV_j  = b3_j M1_j
//   rb4 = -a4 delta ;  bb3 = rowdot(V, M1) + E^T rb4 ;  cb3 = -b3^2 bb3
//   ab3 = E cb3 ; rb3 = -a3^2 ab3 ;  bb2 = E^T rb3 ; cb2 = -b2^2 bb2
//   ab2 = E cb2 ; rb2 = -a2^2 ab2 ;  bb1 = E^T rb2 ; cb1 = -b1^2 bb1
//   ab1 = E cb1 ; rb1 = -a1^2 ab1
//   dE_ij = a4_i b3_j G_ij + rb1_i + rb2_i b1_j + rb3_i b2_j + rb4_i b3_j + a1_i cb1_j + a2_i cb2_j + a3_i cb3_j
//   dS_ij = scale E_ij dE_ij          (P is invariant to a rescaling of a row of E, so the softmax row term vanishes)
//   dQ = dS K ,  dK = dS^T Q.
// Six matrix-vector passes with the one bf16 matrix in shared memory replace the seven N x N sweeps over a gradient and
// a probability matrix of the step-by-step backward, and all five matrix products run on the tensor cores: shared memory
// holds P' (so P'^T dO needs no scaled copy of dO; E x = (P' x) / a4 and E^T y = P'^T (y / a4)), G stays in tensor
// memory as fp32 until the assembly of dS, which overwrites P' in place as the operand of the last two products.
// ================================================================================================================
constexpr int SKB_THREADS = 32 * 9;
constexpr int SKB_NVEC = 9;       // b1 b2 b3 | cb1 cb2 cb3 (per key column) | y (per query row) | column-sum partials [2][256]

struct SkTcBwdParams {
  int B, N, H, NP, items;
  float scale, scale_log2e;
  const bf16* o;       // forward output [B, N, H*64]
  const bf16* dout;    // [B, N, H*64]
  const float* stats;  // [B, H, 8, N]
  bf16* dqkv;          // [B, N, 3, H, 64]
};

__global__ void __launch_bounds__(SKB_THREADS, 1)
sinkhorn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tm128, const __grid_constant__ CUtensorMap tm16,
                       const __grid_constant__ CUtensorMap td128, const __grid_constant__ CUtensorMap td16, const SkTcBwdParams p) {
  pdl_launch_dependents();   // the next kernel in the stream (a PDL-launched GEMM) may start its prologue
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const int NP = p.NP, N = p.N, H = p.H;
  const int TB = NP * 128;                                    // bytes of one [NP rows][64 bf16] tile (NP % 8 == 0: 1024-aligned)
  // P' / dS: 4 key chunks of [NP rows][128 B].  A row tile of 128 read from row 128 runs past row NP into the next
  // buffer: those operand rows only produce output rows that nobody reads.
  const uint32_t sE = sbase, sQ = sE + 4 * TB, sK = sQ + TB, sV = sK + TB, sD = sV + TB;
  const uint32_t sVec = sD + TB;
  const uint32_t bar0 = sVec + SKB_NVEC * SKT_ROWS * 4;
  const uint32_t bar_ld = bar0, bar_mma = bar0 + 8, bar_go = bar0 + 16;
  float* vec = reinterpret_cast<float*>(smem + (sVec - sbase));
  float* vb1 = vec, *vb2 = vec + 256, *vb3 = vec + 512, *vc1 = vec + 768, *vc2 = vec + 1024, *vc3 = vec + 1280, *tmp = vec + 1536, *vpart = vec + 1792;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + (bar0 - sbase) + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 8) {
    if (elect_one()) {
      tma_prefetch_desc(&tm128); tma_prefetch_desc(&tm16); tma_prefetch_desc(&td128); tma_prefetch_desc(&td16);
      mbar_init(bar_ld, 1);
      mbar_init(bar_mma, 1);
      mbar_init(bar_go, 8);       // the 8 arithmetic warps hand a phase back to the control warp
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t T = *tmem_ptr_smem;
  const int my_items = ((int)blockIdx.x < p.items) ? (p.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int ksteps_n = NP / 16;
  const int ntiles = (N + 127) / 128;
  // tensor memory: S, later G: row tile t at columns [t * NP, (t + 1) * NP) ; M1, dQ, dK in the columns those leave free
  const uint32_t T_M = 0;                         // M1 tile t at T_M + t * 64 (while S is dead and G not yet issued)

  if (warp == 8) {
    // ================================ TMA + MMA issue (one lane) ================================
    if (elect_one()) {
      const uint64_t dk = make_smem_desc_sw128(0, 16, 1024);                     // K-major operand
      const uint64_t dmn = make_smem_desc_sw128(0, (uint32_t)TB, 1024);          // MN-major operand, 64-element chunks TB apart
      const uint32_t id_s = make_idesc(1u, 0u, 0u, 128u, (uint32_t)NP);          // S, G: A K-major, B K-major, N = NP
      const uint32_t id_tn = make_idesc(1u, 1u, 1u, 128u, 64u);                  // A^T B: A MN-major, B MN-major, N = 64
      const uint32_t id_nn = make_idesc(1u, 0u, 1u, 128u, 64u);                  // A B:   A K-major,  B MN-major, N = 64
      auto load_rows = [&](const CUtensorMap* t128, const CUtensorMap* t16, uint32_t dst, int c1, int nrows, int b) {
        int r = 0;
        for (; r + 128 <= nrows; r += 128) tma_load_4d(dst + r * 128, t128, bar_ld, 0, c1, r, b);
        for (; r < nrows; r += 16) tma_load_4d(dst + r * 128, t16, bar_ld, 0, c1, r, b);
      };
      uint32_t go = 0, mm = 0;     // phase counters of bar_go / bar_mma
      auto wait_go = [&]() { mbar_wait(bar_go, go & 1, 60); ++go; tc_fence_after(); };
      auto commit_wait = [&]() { umma_commit(bar_mma); mbar_wait(bar_mma, mm & 1, 61); ++mm; };
      for (int li = 0; li < my_items; ++li) {
        const int item = (int)blockIdx.x + li * (int)gridDim.x;
        const int b = item / H, h = item % H;
        if (li > 0) wait_go();                                   // previous item: dQ / dK read out of tensor memory
        mbar_arrive_expect_tx(bar_ld, 4 * TB);
        load_rows(&tm128, &tm16, sQ, 0 * H + h, NP, b);
        load_rows(&tm128, &tm16, sK, 1 * H + h, NP, b);
        load_rows(&tm128, &tm16, sV, 2 * H + h, NP, b);
        load_rows(&td128, &td16, sD, h, NP, b);
        mbar_wait(bar_ld, li & 1, 62);
        tc_fence_after();
        // ---- S = Q K^T
        for (int t = 0; t < ntiles; ++t)
          for (int ks = 0; ks < 4; ++ks)
            umma_bf16(T + t * NP, dk + (uint64_t)((sQ + t * 16384 + ks * 32) >> 4), dk + (uint64_t)((sK + ks * 32) >> 4), id_s, ks > 0);
        commit_wait();
        wait_go();                                               // P' is in shared memory, S has been read
        // ---- M1 = P'^T dO : M = keys (tile t = key chunks 2t, 2t+1), K = queries, N = 64
        for (int t = 0; t < ntiles; ++t)
          for (int ks = 0; ks < ksteps_n; ++ks)
            umma_bf16(T + T_M + t * 64, dmn + (uint64_t)((sE + 2 * t * TB + ks * 2048) >> 4), dmn + (uint64_t)((sD + ks * 2048) >> 4),
                      id_tn, ks > 0);
        commit_wait();
        wait_go();                                               // M1 has been read
        // ---- G = dO V^T
        for (int t = 0; t < ntiles; ++t)
          for (int ks = 0; ks < 4; ++ks)
            umma_bf16(T + t * NP, dk + (uint64_t)((sD + t * 16384 + ks * 32) >> 4), dk + (uint64_t)((sV + ks * 32) >> 4), id_s, ks > 0);
        commit_wait();
        wait_go();                                               // dS is in shared memory (over P'), G has been read
        // ---- dQ = dS K (tiles at columns 0, 64) ; dK = dS^T Q (tiles at columns 128, 192)
        for (int t = 0; t < ntiles; ++t)
          for (int ks = 0; ks < ksteps_n; ++ks)
            umma_bf16(T + t * 64, dk + (uint64_t)((sE + (ks >> 2) * TB + t * 16384 + (ks & 3) * 32) >> 4),
                      dmn + (uint64_t)((sK + ks * 2048) >> 4), id_nn, ks > 0);
        for (int t = 0; t < ntiles; ++t)
          for (int ks = 0; ks < ksteps_n; ++ks)
            umma_bf16(T + 128 + t * 64, dmn + (uint64_t)((sE + 2 * t * TB + ks * 2048) >> 4), dmn + (uint64_t)((sQ + ks * 2048) >> 4),
                      id_tn, ks > 0);
        commit_wait();
      }
    }
    __syncwarp();
  } else {
    // ================================ arithmetic warps: thread = row / column r =================
    const int tile = warp >> 2, q = warp & 3;
    const int r = tile * 128 + q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int nch = NP / 16;
    const uint64_t c2 = f2_pack(p.scale_log2e, p.scale_log2e);
    const bool in_np = r < NP, in_n = r < N;
    const uint32_t e_row = sE + r * 128;
    uint32_t mm = 0;
    auto wait_mma = [&]() { mbar_wait(bar_mma, mm & 1, 70); ++mm; tc_fence_after(); };
    auto hand_back = [&]() { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(bar_go); };
    // y_i = sum_j P'_ij x_j (thread = row; x from shared memory)
    auto matvec_row = [&](const float* x) {
      float acc = 0.f;
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j0 = c * 64 + u * 8;
          if (j0 < NP) {
            const uint4 w = lds128(e_row + c * TB + ((u ^ (r & 7)) << 4));
            const float4 b0 = lds128f(smem_u32(x + j0)), b1 = lds128f(smem_u32(x + j0 + 4));
            const float2 e0 = unpack_bf16(w.x), e1 = unpack_bf16(w.y), e2 = unpack_bf16(w.z), e3 = unpack_bf16(w.w);
            acc = fmaf(e0.x, b0.x, acc); acc = fmaf(e0.y, b0.y, acc); acc = fmaf(e1.x, b0.z, acc); acc = fmaf(e1.y, b0.w, acc);
            acc = fmaf(e2.x, b1.x, acc); acc = fmaf(e2.y, b1.y, acc); acc = fmaf(e3.x, b1.z, acc); acc = fmaf(e3.y, b1.w, acc);
          }
        }
      }
      return acc;
    };
    const int tid = threadIdx.x;
    // y_j = sum_i P'_ij x_i (all 256 threads call it; includes one barrier round)
    auto matvec_col = [&](const float* x) { return skt_colsum(sE, (uint32_t)TB, N, NP, smem_u32(x), vpart, tid, r); };
    for (int li = 0; li < my_items; ++li) {
      const int item = (int)blockIdx.x + li * (int)gridDim.x;
      const int b = item / H, h = item % H;
      const float* st = p.stats + (long long)item * 8 * N;
      // ---- forward vectors from the stashed step sums; delta_i = <dO_i, O_i>
      float a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, lse = 0.f, delta = 0.f;
      if (in_n) {
        lse = st[r];
        a1 = 1.f / st[1 * N + r]; a2 = a1 / st[3 * N + r]; a3 = a2 / st[5 * N + r]; a4 = a3 / st[7 * N + r];
        const float bb1 = 1.f / st[2 * N + r], bb2 = bb1 / st[4 * N + r], bb3 = bb2 / st[6 * N + r];
        vb1[r] = bb1; vb2[r] = bb2; vb3[r] = bb3;
        const uint4* po = reinterpret_cast<const uint4*>(p.o + ((long long)b * N + r) * (H * 64) + (long long)h * 64);
        const uint4* pd = reinterpret_cast<const uint4*>(p.dout + ((long long)b * N + r) * (H * 64) + (long long)h * 64);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint4 x = __ldg(po + u), y = __ldg(pd + u);
          const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = unpack_bf16(xs[e]), g = unpack_bf16(ys[e]);
            delta = fmaf(f.x, g.x, delta);
            delta = fmaf(f.y, g.y, delta);
          }
        }
      } else {
        vb1[r] = 0.f; vb2[r] = 0.f; vb3[r] = 0.f;
      }
      // ---- P' = a4 o softmax(S) as bf16 into the operand layout
      wait_mma();
      const bool warp_active = tile * 128 + q * 32 < NP;      // warp-uniform: tcgen05.ld is a warp-collective instruction
      if (warp_active) {
        const uint32_t T_S = T + tile * NP + lane_addr;
        const float noff = in_n ? (-lse * 1.4426950408889634f + __log2f(a4)) : 0.f;
        const uint64_t noff2 = f2_pack(noff, noff);
        for (int c0 = 0; c0 < 16; c0 += 4) {
          uint32_t v[4][16];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (c0 + k < nch) tmem_ld_32x16(T_S + (c0 + k) * 16, v[k]);
          tmem_wait_ld();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int ch = c0 + k, j0 = ch * 16;
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              float e0 = 0.f, e1 = 0.f;
              if (ch < nch && in_n) {
                float x0, x1;
                f2_unpack(f2_fma(f2_pack(__uint_as_float(v[k][j]), __uint_as_float(v[k][j + 1])), c2, noff2), x0, x1);
                e0 = j0 + j < N ? ex2f(x0) : 0.f;
                e1 = j0 + j + 1 < N ? ex2f(x1) : 0.f;
              }
              pk[j >> 1] = pack_bf16(e0, e1);
            }
            const uint32_t base = e_row + (uint32_t)(j0 >> 6) * TB;
            const int u = (j0 & 63) >> 3;
            if (in_np) {                                      // rows beyond NP belong to the next chunk
              sts128(base + (((u) ^ (r & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
              sts128(base + (((u + 1) ^ (r & 7)) << 4), pk[4], pk[5], pk[6], pk[7]);
            }
          }
        }
      }
      fence_async_smem();
      hand_back();                                            // -> M1 = P'^T dO
      // ---- dV_j = b3_j M1_j ; bb3 (direct part) = <V_j, M1_j>      (thread = key row j = r)
      wait_mma();
      float bb3 = 0.f;
      {
        const uint32_t T_M1 = T + T_M + tile * 64 + lane_addr;
        const float b3 = in_n ? vb3[r] : 0.f;
        bf16* dv = p.dqkv + ((long long)b * N + r) * (3 * H * 64) + (long long)(2 * H + h) * 64;
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t v[2][16];
          tmem_ld_32x16(T_M1 + c0, v[0]);
          tmem_ld_32x16(T_M1 + c0 + 16, v[1]);
          tmem_wait_ld();
          if (in_n) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              // V row r: 16-byte unit (c0 + 16 k) / 8 and the next one, swizzled by the row
              const int u0 = (c0 + 16 * k) >> 3;
              const uint4 va = lds128(sV + r * 128 + (((u0) ^ (r & 7)) << 4)), vb = lds128(sV + r * 128 + (((u0 + 1) ^ (r & 7)) << 4));
              const uint32_t vs[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
              uint32_t w[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float m0 = __uint_as_float(v[k][2 * e]), m1 = __uint_as_float(v[k][2 * e + 1]);
                const float2 f = unpack_bf16(vs[e]);
                bb3 = fmaf(f.x, m0, bb3);
                bb3 = fmaf(f.y, m1, bb3);
                w[e] = pack_bf16(m0 * b3, m1 * b3);
              }
              *reinterpret_cast<uint4*>(dv + c0 + 16 * k) = make_uint4(w[0], w[1], w[2], w[3]);
              *reinterpret_cast<uint4*>(dv + c0 + 16 * k + 8) = make_uint4(w[4], w[5], w[6], w[7]);
            }
          }
        }
      }
      hand_back();                                            // -> G = dO V^T (runs under the vector chain)
      // ---- the chain of six matrix-vector passes.  Shared memory holds P' = a4 o E:  E x = (P' x) / a4 ,  E^T y = P'^T (y / a4)
      const float inv_a4 = in_n ? 1.f / a4 : 0.f;
      tmp[r] = in_n ? -delta : 0.f;                           // rb4 / a4   (rb4 = -a4^2 (delta / a4))
      skt_sync();
      {                                                       // bb3 += E^T rb4 ; cb3 = -b3^2 bb3
        const float t = matvec_col(tmp);
        const float b3 = vb3[r];
        vc3[r] = in_n ? -b3 * b3 * (bb3 + t) : 0.f;
      }
      skt_sync();
      float rb3, rb2, rb1;
      {                                                       // ab3 = E cb3 ; rb3 = -a3^2 ab3
        const float t = in_np ? matvec_row(vc3) : 0.f;
        rb3 = -a3 * a3 * t * inv_a4;
        tmp[r] = rb3 * inv_a4;
      }
      skt_sync();
      {                                                       // bb2 = E^T rb3 ; cb2 = -b2^2 bb2
        const float t = matvec_col(tmp);
        const float b2 = vb2[r];
        vc2[r] = in_n ? -b2 * b2 * t : 0.f;
      }
      skt_sync();
      {                                                       // ab2 = E cb2 ; rb2 = -a2^2 ab2
        const float t = in_np ? matvec_row(vc2) : 0.f;
        rb2 = -a2 * a2 * t * inv_a4;
        tmp[r] = rb2 * inv_a4;
      }
      skt_sync();
      {                                                       // bb1 = E^T rb2 ; cb1 = -b1^2 bb1
        const float t = matvec_col(tmp);
        const float b1 = vb1[r];
        vc1[r] = in_n ? -b1 * b1 * t : 0.f;
      }
      skt_sync();
      {                                                       // ab1 = E cb1 ; rb1 = -a1^2 ab1
        const float t = in_np ? matvec_row(vc1) : 0.f;
        rb1 = -a1 * a1 * t * inv_a4;
      }
      // ---- dS_ij = scale E_ij dE_ij = P'_ij k0 dE_ij (k0 = scale / a4), in place over P'
      wait_mma();                                             // G
      if (warp_active) {
        const uint32_t T_G = T + tile * NP + lane_addr;
        const float k0 = p.scale * inv_a4;
        const uint64_t R1 = f2_pack(k0 * rb1, k0 * rb1), R2 = f2_pack(k0 * rb2, k0 * rb2), R3 = f2_pack(k0 * rb3, k0 * rb3);
        const float r4 = -k0 * a4 * delta;
        const uint64_t R4 = f2_pack(r4, r4);
        const uint64_t A1 = f2_pack(k0 * a1, k0 * a1), A2 = f2_pack(k0 * a2, k0 * a2), A3 = f2_pack(k0 * a3, k0 * a3);
        const uint64_t SC = f2_pack(p.scale, p.scale);       // k0 a4
        for (int ch = 0; ch < nch; ++ch) {
          uint32_t v[16];
          tmem_ld_32x16(T_G + ch * 16, v);
          const int j0 = ch * 16;
          const uint32_t base = e_row + (uint32_t)(j0 >> 6) * TB;
          const int u = (j0 & 63) >> 3;
          uint4 w0 = make_uint4(0, 0, 0, 0), w1 = w0;
          if (in_np) { w0 = lds128(base + (((u) ^ (r & 7)) << 4)); w1 = lds128(base + (((u + 1) ^ (r & 7)) << 4)); }
          const uint32_t es[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
          tmem_wait_ld();
          uint32_t o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int j = j0 + 2 * e;
            const float2 pb1 = *reinterpret_cast<const float2*>(vb1 + j), pb2 = *reinterpret_cast<const float2*>(vb2 + j);
            const float2 pb3 = *reinterpret_cast<const float2*>(vb3 + j), pc1 = *reinterpret_cast<const float2*>(vc1 + j);
            const float2 pc2 = *reinterpret_cast<const float2*>(vc2 + j), pc3 = *reinterpret_cast<const float2*>(vc3 + j);
            const uint64_t B3 = f2_pack(pb3.x, pb3.y);
            uint64_t d = f2_fma(f2_mul(SC, B3), f2_pack(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1])), R1);
            d = f2_fma(R2, f2_pack(pb1.x, pb1.y), d);
            d = f2_fma(R3, f2_pack(pb2.x, pb2.y), d);
            d = f2_fma(R4, B3, d);
            d = f2_fma(A1, f2_pack(pc1.x, pc1.y), d);
            d = f2_fma(A2, f2_pack(pc2.x, pc2.y), d);
            d = f2_fma(A3, f2_pack(pc3.x, pc3.y), d);
            const float2 pe = unpack_bf16(es[e]);
            float x0, x1;
            f2_unpack(f2_mul(d, f2_pack(pe.x, pe.y)), x0, x1);
            o[e] = in_n ? pack_bf16(x0, x1) : 0u;
          }
          if (in_np) {
            sts128(base + (((u) ^ (r & 7)) << 4), o[0], o[1], o[2], o[3]);
            sts128(base + (((u + 1) ^ (r & 7)) << 4), o[4], o[5], o[6], o[7]);
          }
        }
      }
      fence_async_smem();
      hand_back();                                            // -> dQ = dS K, dK = dS^T Q
      // ---- epilogue: dQ row r, dK row r
      wait_mma();
      {
        bf16* dq = p.dqkv + ((long long)b * N + r) * (3 * H * 64) + (long long)h * 64;
        bf16* dkp = dq + (long long)H * 64;
#pragma unroll
        for (int which = 0; which < 2; ++which) {
          const uint32_t T_O = T + which * 128 + tile * 64 + lane_addr;
          bf16* dst = which == 0 ? dq : dkp;
#pragma unroll
          for (int c0 = 0; c0 < 64; c0 += 32) {
            uint32_t v[2][16];
            tmem_ld_32x16(T_O + c0, v[0]);
            tmem_ld_32x16(T_O + c0 + 16, v[1]);
            tmem_wait_ld();
            if (in_n) {
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                uint32_t w[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) w[e] = pack_bf16(__uint_as_float(v[k][2 * e]), __uint_as_float(v[k][2 * e + 1]));
                *reinterpret_cast<uint4*>(dst + c0 + 16 * k) = make_uint4(w[0], w[1], w[2], w[3]);
                *reinterpret_cast<uint4*>(dst + c0 + 16 * k + 8) = make_uint4(w[4], w[5], w[6], w[7]);
              }
            }
          }
        }
      }
      hand_back();                                            // tensor memory and the operand tiles are free for the next item
      skt_sync();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(T, 512);
}

bool sinkhorn_tc_supported(int N, int dh, int dtype) { return dtype == NRV_BF16 && dh == 64 && N >= 1 && N <= SKT_MAXN; }

static int skt_smem_bytes(int NP) {
  const int kva = (NP * 128 + 1023) & ~1023;
  return 4 * SKT_CHUNK + SKT_ROWS * 128 + 2 * kva + 4 * SKT_ROWS * 4 + 128 + 1024;
}

int sinkhorn_fwd_tc(const void* qkv, void* out, float* stats, int B, int N, int H, int dh, float scale, cudaStream_t st) {
  NRV_REQUIRE(sinkhorn_tc_supported(N, dh, NRV_BF16), "tcgen05 Sinkhorn attention: unsupported shape N=%d dh=%d", N, dh);
  NRV_REQUIRE(stats != nullptr, "Sinkhorn attention needs the [B,H,8,N] fp32 statistics buffer");
  NRV_REQUIRE(((uintptr_t)qkv % 16) == 0 && ((uintptr_t)out % 16) == 0, "tcgen05 Sinkhorn attention: 16-byte alignment");
  SkTcParams p{};
  p.B = B; p.N = N; p.H = H; p.NP = (N + 15) / 16 * 16; p.items = B * H;
  p.scale = scale; p.scale_log2e = scale * 1.4426950408889634f;
  p.out = (bf16*)out; p.stats = stats;
  const uint64_t dims[4] = {(uint64_t)dh, (uint64_t)3 * H, (uint64_t)N, (uint64_t)B};
  const uint64_t strides[3] = {(uint64_t)dh * 2, (uint64_t)3 * H * dh * 2, (uint64_t)N * 3 * H * dh * 2};
  const uint32_t box128[4] = {64, 1, 128, 1}, box16[4] = {64, 1, 16, 1};
  CUtensorMap t128, t16;
  int rc = encode_tmap_4d(&t128, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, qkv, dims, strides, box128, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = encode_tmap_4d(&t16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, qkv, dims, strides, box16, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  const int smem = skt_smem_bytes(p.NP);
  NRV_REQUIRE(smem <= 227 * 1024, "tcgen05 Sinkhorn attention: %d bytes of shared memory", smem);
  NRV_CUDA(cudaFuncSetAttribute(sinkhorn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = p.items < num_sms() ? p.items : num_sms();
  sinkhorn_tc_fwd_kernel<<<grid, SKT_THREADS, smem, st>>>(t128, t16, p);
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

int sinkhorn_bwd_tc(const void* qkv, const void* out, const void* dout, const float* stats, void* dqkv, int B, int N, int H, int dh,
                    float scale, cudaStream_t st) {
  NRV_REQUIRE(sinkhorn_tc_supported(N, dh, NRV_BF16), "tcgen05 Sinkhorn attention: unsupported shape N=%d dh=%d", N, dh);
  NRV_REQUIRE(qkv && out && dout && stats && dqkv, "tcgen05 Sinkhorn attention backward: null pointer");
  NRV_REQUIRE(((uintptr_t)qkv % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)dout % 16) == 0 && ((uintptr_t)dqkv % 16) == 0,
              "tcgen05 Sinkhorn attention: 16-byte alignment");
  SkTcBwdParams p{};
  p.B = B; p.N = N; p.H = H; p.NP = (N + 15) / 16 * 16; p.items = B * H;
  p.scale = scale; p.scale_log2e = scale * 1.4426950408889634f;
  p.o = (const bf16*)out; p.dout = (const bf16*)dout; p.stats = stats; p.dqkv = (bf16*)dqkv;
  const uint32_t box128[4] = {64, 1, 128, 1}, box16[4] = {64, 1, 16, 1};
  CUtensorMap t128, t16, d128, d16;
  {
    const uint64_t dims[4] = {(uint64_t)dh, (uint64_t)3 * H, (uint64_t)N, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)dh * 2, (uint64_t)3 * H * dh * 2, (uint64_t)N * 3 * H * dh * 2};
    int rc = encode_tmap_4d(&t128, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, qkv, dims, strides, box128, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = encode_tmap_4d(&t16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, qkv, dims, strides, box16, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)dh, (uint64_t)H, (uint64_t)N, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)dh * 2, (uint64_t)H * dh * 2, (uint64_t)N * H * dh * 2};
    int rc = encode_tmap_4d(&d128, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dout, dims, strides, box128, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = encode_tmap_4d(&d16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dout, dims, strides, box16, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  // a row tile of 128 read from row 128 of the last operand tile runs (256 - NP) * 128 <= 6144 bytes into the vectors behind it
  // (shorter sequences: the row tiles reach further than the tile plus the vectors; allocated explicitly)
  const int tb = p.NP * 128;
  const int after_last_tile = SKB_NVEC * SKT_ROWS * 4 + 128;
  const int tile_reach = ((N + 127) / 128) * 16384;     // bytes a K-major A operand reads from the start of its buffer
  const int smem = 7 * tb + (tb + after_last_tile > tile_reach ? tb + after_last_tile : tile_reach) + 1024;
  NRV_REQUIRE(smem <= 227 * 1024, "tcgen05 Sinkhorn attention backward: %d bytes of shared memory", smem);
  NRV_CUDA(cudaFuncSetAttribute(sinkhorn_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = p.items < num_sms() ? p.items : num_sms();
  sinkhorn_tc_bwd_kernel<<<grid, SKB_THREADS, smem, st>>>(t128, t16, d128, d16, p);
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

}  // namespace nrv
