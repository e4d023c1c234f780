// tcgen05 attention forward for the shapes attention_fwd2.cu does not cover: head dims 16..128 (multiples of 16)
// and up to 384 tokens -- ViT-H/14 (dh = 80, N = 257) was on the CUDA-core fallback before.
//   dots = q k^T * scale ; attn = softmax(dots) ; out = attn v          (simple_vit.py:70-75 ; utils.py:207-232)
//
// One pipeline per SM (the score row alone takes up to 384 of the 512 TMEM columns): 4 softmax warps (thread =
// score row) + 1 control warp.  Same building blocks as attention_fwd2.cu, generalised:
//   * operands come through ONE 4-D tensor map over the packed projection output [B, N, 3H, dh]; a head is loaded as
//     ceil(dh/64) column chunks of 64 -- columns beyond dh are out of bounds in the innermost dimension, so TMA
//     zero-fills them and every chunk is a regular 128-byte-row SWIZZLE_128B tile
//   * S = Q K^T: dh/16 K-steps, the keys in slices of at most 256 (UMMA N limit), straight into TMEM columns [0, NP)
//   * softmax: two passes over the TMEM row in rounds of 64 columns (max, then exp2 + row sum), P written back as
//     bf16 over the consumed scores
//   * O = P V: A operand from tensor memory (TS form), V re-read MN-major across its chunks (N = chunks x 64;
//     the zero columns cost a few idle MMA columns and nothing else), O in TMEM columns [384, 384 + chunks x 64)
//   * each thread stores its output row (dh bf16, contiguous) directly
// K/V of an item and the Q tiles are single-buffered: this kernel serves inference-sized problems, the training
// path (N <= 208, dh = 64) has its own double-pipelined kernel.
#include "common.cuh"
#include "nrvit_internal.h"

namespace nrv {

constexpr int FB_THREADS = 32 * 5;
constexpr int FB_T_O = 384;            // first TMEM column of O

struct FwdBigParams {
  int B, N, H, dh, NP, CH, tiles, items;
  int kvb;                 // bytes of one K / V chunk: NP * 128
  float scale, scale_log2e;
  bf16* out;               // [B, N, H*dh]
  float* lse;              // [B, H, N] or null
  AttnDrop dr;             // dropout on the probabilities (dr.p = 0: none): out = (P o M) V, row sum over the unmasked P
  uint32_t dthresh;        // p * 2^24
  float dscale;            // 1 / (1 - p)
};

// f2_exp16 (common.cuh) with the dropout mask: `keep` bit t = column 16c + t survives.  The row sum takes the unmasked
// probabilities; the bf16 pairs that go back to tensor memory (the A operand of P V) are zeroed where dropped, and the
// 1 / (1 - p) rides on the row's final 1 / sum.
__device__ __forceinline__ void f2_exp16_drop(const uint32_t (&v)[16], int c, int N, uint64_t c2, uint64_t noff2,
                                              uint64_t& sum2, uint32_t t_p, uint32_t keep) {
  uint32_t pk[8];
  const int c0 = c * 16;
  const bool full = c0 + 16 <= N;
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    const uint64_t x2 = f2_fma(f2_pack(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), c2, noff2);
    float x0, x1;
    f2_unpack(x2, x0, x1);
    float e0 = ex2f(x0), e1 = ex2f(x1);
    if (!full) {
      if (c0 + j >= N) e0 = 0.f;
      if (c0 + j + 1 >= N) e1 = 0.f;
    }
    sum2 = f2_add(sum2, f2_pack(e0, e1));
    const uint32_t m = (((keep >> j) & 1u) ? 0x0000ffffu : 0u) | (((keep >> (j + 1)) & 1u) ? 0xffff0000u : 0u);
    pk[j >> 1] = pack_bf16(e0, e1) & m;
  }
  tmem_st_32x8(t_p + c * 8, pk);
}

__global__ void __launch_bounds__(FB_THREADS, 1)
attn_fwd_big_kernel(const __grid_constant__ CUtensorMap tm128, const __grid_constant__ CUtensorMap tm16, const FwdBigParams p) {
  pdl_launch_dependents();   // the next kernel in the stream (a PDL-launched GEMM) may start its prologue
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const int NP = p.NP, N = p.N, H = p.H, CH = p.CH, KVB = p.kvb;
  const uint32_t sQ = sbase, sK = sQ + CH * 16384, sV = sK + CH * KVB;
  const uint32_t bar0 = sV + CH * KVB;
  const uint32_t bar_q = bar0, bar_kv = bar0 + 8, bar_s = bar0 + 16, bar_p = bar0 + 24, bar_o = bar0 + 32, bar_free = bar0 + 40;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + (bar0 - sbase) + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 4) {
    if (elect_one()) {
      tma_prefetch_desc(&tm128);
      tma_prefetch_desc(&tm16);
      mbar_init(bar_q, 1);
      mbar_init(bar_kv, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_p, 4);
      mbar_init(bar_o, 1);
      mbar_init(bar_free, 4);
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
  const int total_tiles = my_items * p.tiles;
  const int ksteps_d = p.dh / 16, ksteps_n = NP / 16;

  if (warp == 4) {
    // ================================ TMA + MMA issue (one lane) ================================
    if (elect_one() && total_tiles > 0) {
      const uint64_t dfix = make_smem_desc_sw128(0, 16, 1024);
      const uint64_t dfix_v = make_smem_desc_sw128(0, (uint32_t)KVB, 1024);     // MN-major B across the V chunks
      const uint32_t idesc_o = make_idesc(1u, 0u, 1u, 128u, (uint32_t)(CH * 64));
      auto load_rows = [&](uint32_t dst, uint32_t bar, int which, int h, int row0, int nrows, int b) {
        // rows [row0, row0 + nrows) of q (which = 0) / k (1) / v (2) of head h, all column chunks; nrows = 128a + 16c
        for (int c = 0; c < CH; ++c) {
          int r = 0;
          for (; r + 128 <= nrows; r += 128)
            tma_load_4d(dst + c * (which == 0 ? 16384 : KVB) + r * 128, &tm128, bar, c * 64, which * H + h, row0 + r, b);
          for (; r < nrows; r += 16)
            tma_load_4d(dst + c * (which == 0 ? 16384 : KVB) + r * 128, &tm16, bar, c * 64, which * H + h, row0 + r, b);
        }
      };
      // (running tile counters and one division pair per item: no integer division per tile on the issuing thread)
      auto issue_q = [&](int li_n, int t_n) {
        const int item = (int)blockIdx.x + li_n * (int)gridDim.x;
        const int b_n = item / H;
        mbar_arrive_expect_tx(bar_q, CH * 16384);
        load_rows(sQ, bar_q, 0, item - b_n * H, t_n * 128, 128, b_n);
      };
      issue_q(0, 0);
      for (int g = 0, li = 0, t = 0; g < total_tiles; ++g, (++t == p.tiles ? (t = 0, ++li) : 0)) {
        const uint32_t ph = g & 1;
        if (t == 0) {
          // the previous item's last P V has retired (bar_o was waited below): K / V may be overwritten
          const int item = (int)blockIdx.x + li * (int)gridDim.x;
          const int b = item / H, h = item - b * H;
          mbar_arrive_expect_tx(bar_kv, 2 * CH * KVB);
          load_rows(sK, bar_kv, 1, h, 0, NP, b);
          load_rows(sV, bar_kv, 2, h, 0, NP, b);
          mbar_wait(bar_kv, li & 1, 11);
        }
        mbar_wait(bar_q, ph, 10);
        tc_fence_after();
        for (int n0 = 0; n0 < NP; n0 += 256) {               // key slices of at most 256 (UMMA N limit)
          const uint32_t idesc_s = make_idesc(1u, 0u, 0u, 128u, (uint32_t)min(256, NP - n0));
          for (int ks = 0; ks < ksteps_d; ++ks) {
            const uint64_t a = dfix + (uint64_t)((sQ + (ks >> 2) * 16384 + (ks & 3) * 32) >> 4);
            const uint64_t b = dfix + (uint64_t)((sK + (ks >> 2) * KVB + n0 * 128 + (ks & 3) * 32) >> 4);
            umma_bf16(T + n0, a, b, idesc_s, ks > 0);
          }
        }
        umma_commit(bar_s);
        mbar_wait(bar_s, ph, 12);                             // S done: the Q tile may be overwritten
        if (g + 1 < total_tiles) issue_q(t + 1 == p.tiles ? li + 1 : li, t + 1 == p.tiles ? 0 : t + 1);
        mbar_wait(bar_p, ph, 13);                             // P written to TMEM
        if (g > 0) mbar_wait(bar_free, (g - 1) & 1, 14);      // the previous tile's O is in registers
        tc_fence_after();
        for (int ks = 0; ks < ksteps_n; ++ks)
          umma_bf16_ts(T + FB_T_O, T + ks * 8, dfix_v + (uint64_t)((sV + ks * 2048) >> 4), idesc_o, ks > 0);
        umma_commit(bar_o);
        mbar_wait(bar_o, ph, 15);                             // (also frees the score columns and, at t = last, K / V)
      }
    }
    __syncwarp();
  } else {
    // ================================ softmax + epilogue warps =================================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t T_S = T + lane_addr, T_Ov = T + FB_T_O + lane_addr;
    const int nch = NP / 16;
    const uint64_t c2 = f2_pack(p.scale_log2e, p.scale_log2e);
    const long long HD = (long long)H * p.dh;
    int b = 0, h = 0;
    for (int g = 0, li = 0, t = 0; g < total_tiles; ++g, (++t == p.tiles ? (t = 0, ++li) : 0)) {
      if (t == 0) {
        const int item = (int)blockIdx.x + li * (int)gridDim.x;
        b = item / H; h = item - b * H;
      }
      const uint32_t ph = g & 1;
      const int n = t * 128 + r;
      const bool warp_active = t * 128 + q * 32 < N;
      mbar_wait(bar_s, ph, 20);
      tc_fence_after();
      float mx = 0.f, tot = 1.f;
      if (warp_active) {
        // pass 1: row max, in rounds of four 16-column chunks
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
        // pass 2: p = exp2(s*c - mx*c), row sum, bf16 pairs back over the consumed scores (P chunk k sits at columns
        // [8k, 8k+8): inside score chunks <= k, never ahead of the round being read)
        const float noff = -mx * p.scale_log2e;
        const uint64_t noff2 = f2_pack(noff, noff);
        uint64_t sum2 = f2_pack(0.f, 0.f);
        for (int c0 = 0; c0 < nch; c0 += 4) {
          uint32_t v[4][16];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (c0 + k < nch) tmem_ld_32x16(T_S + (c0 + k) * 16, v[k]);
          tmem_wait_ld();
          if (p.dr.p > 0.f) {
            // element ((b*H + h)*N + n)*N + key of the mask stream (attn_keep16: 4-5 Philox calls per 16 keys)
            const unsigned long long erow = (((unsigned long long)b * H + h) * N + (unsigned long long)n) * N;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (c0 + k < nch)
                f2_exp16_drop(v[k], c0 + k, N, c2, noff2, sum2, T_S, attn_keep16(p.dr, p.dthresh, erow + (c0 + k) * 16));
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (c0 + k < nch) f2_exp16(v[k], c0 + k, N, c2, noff2, sum2, T_S);
          }
        }
        tmem_wait_st();
        float s0, s1;
        f2_unpack(sum2, s0, s1);
        tot = s0 + s1;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
      const float inv = __fdividef(p.dr.p > 0.f ? p.dscale : 1.f, tot);
      if (warp_active && n < N && p.lse) p.lse[((long long)b * H + h) * N + n] = mx * p.scale + __logf(tot);

      mbar_wait(bar_o, ph, 21);
      tc_fence_after();
      // epilogue: dh columns of O in chunks of 16, scaled, stored as this thread's contiguous output row
      bf16* dst = p.out + ((long long)b * N + n) * HD + (long long)h * p.dh;
      for (int c0 = 0; c0 < p.dh; c0 += 32) {
        uint32_t v[2][16];
        const bool two = c0 + 16 < p.dh;
        if (warp_active) {
          tmem_ld_32x16(T_Ov + c0, v[0]);
          if (two) tmem_ld_32x16(T_Ov + c0 + 16, v[1]);
          tmem_wait_ld();
          if (n < N) {
#pragma unroll
            for (int k = 0; k < 2; ++k)
              if (k == 0 || two) {
                uint32_t w[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) w[e] = pack_bf16(__uint_as_float(v[k][2 * e]) * inv, __uint_as_float(v[k][2 * e + 1]) * inv);
                *reinterpret_cast<uint4*>(dst + c0 + 16 * k) = make_uint4(w[0], w[1], w[2], w[3]);
                *reinterpret_cast<uint4*>(dst + c0 + 16 * k + 8) = make_uint4(w[4], w[5], w[6], w[7]);
              }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_free);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(T, 512);
}

bool attn_big_supported(int N, int dh, int dtype) {
  if (dtype != NRV_BF16 || dh % 16 != 0 || dh < 16 || dh > 128 || N < 1) return false;
  const int NP = (N + 15) / 16 * 16, CH = (dh + 63) / 64;
  if (NP > FB_T_O) return false;
  return CH * 16384 + 2 * CH * NP * 128 + 128 + 1024 <= 227 * 1024;
}

int attn_fwd_big(const void* qkv, void* out, float* lse, int B, int N, int H, int dh, float scale, cudaStream_t st,
                 float p_drop, unsigned long long seed, int layer) {
  NRV_REQUIRE(attn_big_supported(N, dh, NRV_BF16), "tcgen05 attention (general): unsupported shape N=%d dh=%d", N, dh);
  NRV_REQUIRE(((uintptr_t)qkv % 16) == 0 && ((uintptr_t)out % 16) == 0, "tcgen05 attention: 16-byte alignment");
  FwdBigParams p{};
  p.B = B; p.N = N; p.H = H; p.dh = dh; p.NP = (N + 15) / 16 * 16; p.CH = (dh + 63) / 64;
  p.tiles = (N + 127) / 128; p.items = B * H; p.kvb = p.NP * 128;
  p.scale = scale; p.scale_log2e = scale * 1.4426950408889634f;
  p.out = (bf16*)out; p.lse = lse;
  p.dr = attn_make_drop(p_drop, seed, layer);
  p.dthresh = (uint32_t)(p_drop * 16777216.0f);
  p.dscale = 1.f / (1.f - p_drop);
  const uint64_t dims[4] = {(uint64_t)dh, (uint64_t)3 * H, (uint64_t)N, (uint64_t)B};
  const uint64_t strides[3] = {(uint64_t)dh * 2, (uint64_t)3 * H * dh * 2, (uint64_t)N * 3 * H * dh * 2};
  const uint32_t box128[4] = {64, 1, 128, 1}, box16[4] = {64, 1, 16, 1};
  CUtensorMap t128, t16;
  int rc = encode_tmap_4d(&t128, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, qkv, dims, strides, box128, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = encode_tmap_4d(&t16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, qkv, dims, strides, box16, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  const int smem = p.CH * 16384 + 2 * p.CH * p.kvb + 128 + 1024;
  NRV_CUDA(cudaFuncSetAttribute(attn_fwd_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = p.items < num_sms() ? p.items : num_sms();
  attn_fwd_big_kernel<<<grid, FB_THREADS, smem, st>>>(t128, t16, p);
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

}  // namespace nrv
