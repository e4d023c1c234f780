// tcgen05 attention backward, second generation: ONE fused pass per (batch, head) item instead of the two
// recompute passes of attention_tc.cu (dQ pass + dK/dV pass).  Autograd of
//   dots = q k^T * scale ; attn = softmax(dots) ; out = attn v          (simple_vit.py:70-75)
// with P recomputed once from the stored log-sum-exp:
//   P = exp(S*scale - lse) ; dP = dO V^T ; dS = P o (dP - delta) ; delta_q = <dO_q, O_q>
//   dV = P^T dO ; dK = scale * dS^T Q ; dQ = scale * dS K
//
// Work decomposition (flash-attention-2 style, transposed): key tiles of 128 rows (j) x query blocks of 64
// columns (i).  Per block the tensor core produces S^T = K_j Q_i^T and dP^T = V_j dO_i^T (key rows in the TMEM
// lanes, queries in the columns), 8 SIMT warps (thread = key row, two warps split the 64 query columns) turn
// them into P^T and dS^T, and three accumulating products consume those:
//   dV_j += P^T dO_i     A operand = P^T, bf16, written back into TENSOR MEMORY over S^T (tcgen05.st, "TS" MMA)
//   dK_j += dS^T Q_i     A operand = dS^T in shared memory, K-major (row = key, 128 B = 64 queries)
//   dQ_I += dS  K_j      the SAME shared-memory tile read MN-major (M = queries); issued once per pair of blocks
// B operands are the TMA-loaded Q / dO / K tiles re-read MN-major, exactly as the forward re-reads V.
// Tensor memory (512 columns): two block buffers {S^T 64 | dP^T 64} so block b+1's scores are computed while
// block b is in the SIMT warps; dV_j 64 ; dK_j 64 ; dQ_0 64 ; dQ_1 64  (N <= 256 tokens).
// exp / dS arithmetic is packed f32x2; the per-query vectors (-lse*log2e, -delta) live in shared memory.
// Zero padding does the masking: rows >= N of every tile are zero-filled by TMA (or zeroed once here), the
// padded entries of the vectors are 0, so padded queries give dP = 0, delta = 0 -> dS = 0 and meet dO = 0.
#include "common.cuh"
#include "nrvit_internal.h"

namespace nrv {

constexpr int B2_DH = 64;
constexpr int B2_SIMT_WARPS = 8;
constexpr int B2_THREADS = 32 * (B2_SIMT_WARPS + 1);    // + control warp (TMA + MMA issue)
constexpr int B2_CHUNK = 128 * 128;                     // [128 rows x 64 bf16] swizzled tile, bytes

struct Bwd2Params {
  int B, N, H, NP;       // NP = N rounded up to 16 (<= 256)
  int KT;                // key tiles of 128 rows
  int NQB;               // query blocks of 64 columns (the last one may be 16..64 wide)
  int items;             // B * H
  float scale, scale_log2e;
  const bf16* o;         // [B, N, H*dh]
  const bf16* dout;      // [B, N, H*dh]
  const float* lse;      // [B, H, N]
  bf16* dqkv;            // [B, N, 3, H, dh]
};

struct Bwd2Smem {
  __host__ __device__ static int qd_bytes(int NP) { return (NP * 128 + 1023) & ~1023; }
  __host__ __device__ static int off_k() { return 0; }
  __host__ __device__ static int off_v(int KT) { return KT * B2_CHUNK; }
  __host__ __device__ static int off_q(int KT) { return 2 * KT * B2_CHUNK; }
  __host__ __device__ static int off_do(int KT, int NP) { return off_q(KT) + qd_bytes(NP); }
  __host__ __device__ static int off_ds(int KT, int NP) { return off_do(KT, NP) + qd_bytes(NP); }   // 2 pair buffers x 2 chunks
  __host__ __device__ static int off_vec(int KT, int NP) { return off_ds(KT, NP) + 4 * B2_CHUNK; }  // nlse[256], ndel[256]
  __host__ __device__ static int off_bar(int KT, int NP) { return off_vec(KT, NP) + 2 * 256 * 4; }
  __host__ __device__ static int total(int KT, int NP) { return off_bar(KT, NP) + 8 * 8 + 16 + 1024; }
};

__device__ __forceinline__ float b2_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(B2_THREADS, 1)
attn_bwd2_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                 const Bwd2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const int NP = p.NP, N = p.N, H = p.H, KT = p.KT, NQB = p.NQB;
  const uint32_t sK = sbase + Bwd2Smem::off_k(), sV = sbase + Bwd2Smem::off_v(KT), sQ = sbase + Bwd2Smem::off_q(KT),
                 sdO = sbase + Bwd2Smem::off_do(KT, NP), sdS = sbase + Bwd2Smem::off_ds(KT, NP);
  uint8_t* dS_gen = smem + Bwd2Smem::off_ds(KT, NP);
  float* nlse_s = reinterpret_cast<float*>(smem + Bwd2Smem::off_vec(KT, NP));
  float* ndel_s = nlse_s + 256;
  const uint32_t bar0 = sbase + Bwd2Smem::off_bar(KT, NP);
  const uint32_t bar_load = bar0, bar_s0 = bar0 + 8 /* [2] */, bar_p0 = bar0 + 24 /* [2] */, bar_row = bar0 + 40,
                 bar_accfree = bar0 + 48;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + Bwd2Smem::off_bar(KT, NP) + 64);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int CTRL = B2_SIMT_WARPS;

  // rows NP .. KT*128-1 of the K and V tiles are never written by TMA: zero them once, so that padded key rows
  // give S^T = dP^T = 0 (finite P, dS) and contribute nothing to dQ
  {
    const int pad_rows = KT * 128 - NP;
    for (int i = threadIdx.x; i < pad_rows * 8; i += B2_THREADS) {
      const int off = NP * 128 + i * 16;
      *reinterpret_cast<uint4*>(smem + Bwd2Smem::off_k() + off) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(smem + Bwd2Smem::off_v(KT) + off) = make_uint4(0, 0, 0, 0);
    }
    // the dS pair buffers feed the dQ product with up to 64 stale query columns when a pair is incomplete:
    // start them finite
    for (int i = threadIdx.x; i < 4 * B2_CHUNK / 16; i += B2_THREADS)
      *reinterpret_cast<uint4*>(dS_gen + i * 16) = make_uint4(0, 0, 0, 0);
    fence_async_smem();
  }
  if (warp == CTRL) {
    if (elect_one()) {
      tma_prefetch_desc(&tm_qkv);
      tma_prefetch_desc(&tm_do);
      mbar_init(bar_load, 1);
      mbar_init(bar_s0, 1);
      mbar_init(bar_s0 + 8, 1);
      mbar_init(bar_p0, B2_SIMT_WARPS);
      mbar_init(bar_p0 + 8, B2_SIMT_WARPS);
      mbar_init(bar_row, 1);
      mbar_init(bar_accfree, B2_SIMT_WARPS);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t T = *tmem_ptr_smem;
  const uint32_t T_DV = T + 256, T_DK = T + 320, T_DQ = T + 384;

  const int my_items = ((int)blockIdx.x < p.items) ? (p.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int NB = KT * NQB;                      // blocks per item
  auto nq_of = [&](int i) { return min(64, NP - 64 * i); };   // query columns of block i (multiple of 16)

  if (warp == CTRL) {
    // ================================ TMA + MMA issue (one lane) ================================
    if (elect_one()) {
      const uint32_t idesc_acc = make_idesc(1u, 0u, 1u, 128u, 64u);    // dV / dK: A K-major (TMEM or smem), B MN-major
      const uint32_t idesc_dq = make_idesc(1u, 1u, 1u, 128u, 64u);     // dQ: A MN-major, B MN-major
      auto mma1 = [&](int lb, int gb) {   // S^T and dP^T of local block lb into TMEM buffer gb & 1
        const int j = lb / NQB, i = lb % NQB;
        const uint32_t idesc1 = make_idesc(1u, 0u, 0u, 128u, (uint32_t)nq_of(i));
        const uint32_t tb = T + (uint32_t)(gb & 1) * 128u;
        const uint64_t ak = make_smem_desc_sw128(sK + j * B2_CHUNK, 16, 1024), bq = make_smem_desc_sw128(sQ + i * 64 * 128, 16, 1024);
        const uint64_t av = make_smem_desc_sw128(sV + j * B2_CHUNK, 16, 1024), bd = make_smem_desc_sw128(sdO + i * 64 * 128, 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tb, ak + 2 * k, bq + 2 * k, idesc1, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tb + 64, av + 2 * k, bd + 2 * k, idesc1, k > 0);
        umma_commit(bar_s0 + 8 * (gb & 1));
      };
      int gb = 0, gp = 0, gr = 0;          // running block / pair / row counters (barrier phases)
      for (int li = 0; li < my_items; ++li) {
        const int item = (int)blockIdx.x + li * (int)gridDim.x;
        const int b = item / H, h = item % H;
        // every product of the previous item has retired (its last bar_row): the operand tiles may be overwritten
        if (li > 0) mbar_wait(bar_row, (gr - 1) & 1, 10);
        mbar_arrive_expect_tx(bar_load, 4 * NP * 128);
        tma_load_3d(sK, &tm_qkv, bar_load, (1 * H + h) * B2_DH, 0, b);
        tma_load_3d(sQ, &tm_qkv, bar_load, (0 * H + h) * B2_DH, 0, b);
        tma_load_3d(sV, &tm_qkv, bar_load, (2 * H + h) * B2_DH, 0, b);
        tma_load_3d(sdO, &tm_do, bar_load, h * B2_DH, 0, b);
        mbar_wait(bar_load, li & 1, 11);
        tc_fence_after();
        mma1(0, gb);
        if (NB > 1) mma1(1, gb + 1);
        for (int lb = 0; lb < NB; ++lb) {
          const int j = lb / NQB, i = lb % NQB;
          const int g = gb + lb, u = g & 1;
          const int ks_q = nq_of(i) / 16;
          mbar_wait(bar_p0 + 8 * u, (g >> 1) & 1, 12);      // P^T in TMEM, dS^T in smem
          // the accumulators of the previous row (or item) have been read by the epilogue
          if (i == 0 && gr > 0) mbar_wait(bar_accfree, (gr - 1) & 1, 13);
          tc_fence_after();
          const uint32_t tb = T + (uint32_t)u * 128u;
          const uint32_t ds_chunk = sdS + ((gp & 1) * 2 + (i & 1)) * B2_CHUNK;
          for (int ks = 0; ks < ks_q; ++ks) {               // dV_j += P^T dO_i   (K = queries of the block)
            const uint64_t bd = make_smem_desc_sw128(sdO + i * 64 * 128 + ks * 2048, 16, 1024);
            umma_bf16_ts(T_DV, tb + ks * 16, bd, idesc_acc, (i > 0 || ks > 0) ? 1u : 0u);   // P^T chunk ks sits at column 16 ks
          }
          for (int ks = 0; ks < ks_q; ++ks) {               // dK_j += dS^T Q_i
            const uint64_t ad = make_smem_desc_sw128(ds_chunk + ks * 32, 16, 1024);
            const uint64_t bq = make_smem_desc_sw128(sQ + i * 64 * 128 + ks * 2048, 16, 1024);
            umma_bf16(T_DK, ad, bq, idesc_acc, (i > 0 || ks > 0) ? 1u : 0u);
          }
          if ((i & 1) || i == NQB - 1) {                    // dQ_I += dS K_j over the pair's 128 queries (K = 128 keys)
            const int I = i >> 1;
            for (int ks = 0; ks < 8; ++ks) {
              const uint64_t ad = make_smem_desc_sw128(sdS + (gp & 1) * 2 * B2_CHUNK + ks * 2048, B2_CHUNK, 1024);
              const uint64_t bk = make_smem_desc_sw128(sK + j * B2_CHUNK + ks * 2048, 16, 1024);
              umma_bf16(T_DQ + 64 * I, ad, bk, idesc_dq, (j > 0 || ks > 0) ? 1u : 0u);
            }
            ++gp;
          }
          if (i == NQB - 1) { umma_commit(bar_row); ++gr; }  // dV_j / dK_j complete (and dQ when j is the last row)
          if (lb + 2 < NB) mma1(lb + 2, g + 2);
        }
        gb += NB;
      }
    }
    __syncwarp();
  } else {
    // ================================ SIMT warps: P^T, dS^T, epilogues ==========================
    const int q = warp & 3;                               // TMEM lane quarter
    const int hf = warp >> 2;                             // column half of the block / which accumulator to store
    const int r = q * 32 + lane;                          // key row within the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const long long HD = (long long)H * B2_DH;
    const uint64_t c2 = f2_pack(p.scale_log2e, p.scale_log2e);
    int gb = 0, gp = 0, gr = 0;
    // one epilogue = this warp stores 32 rows x 64 columns of one accumulator
    auto store_acc = [&](uint32_t tsrc, float mul, bf16* dst, bool row_ok) {
      uint32_t v[2][32];
      tmem_ld_32x32(tsrc + lane_addr, v[0]);
      tmem_ld_32x32(tsrc + lane_addr + 32, v[1]);
      tmem_wait_ld();
      if (row_ok) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
              w[k] = pack_bf16(__uint_as_float(v[hh][8 * u + 2 * k]) * mul, __uint_as_float(v[hh][8 * u + 2 * k + 1]) * mul);
            *reinterpret_cast<uint4*>(dst + hh * 32 + u * 8) = make_uint4(w[0], w[1], w[2], w[3]);
          }
      }
    };
    for (int li = 0; li < my_items; ++li) {
      const int item = (int)blockIdx.x + li * (int)gridDim.x;
      const int b = item / H, h = item % H;
      bf16* dq_base = p.dqkv + (long long)b * N * 3 * HD + (long long)h * B2_DH;   // + n*3*HD + which*HD
      auto epilogue_row = [&](int j, bool last) {
        mbar_wait(bar_row, gr & 1, 20);
        tc_fence_after();
        const int n = j * 128 + r;
        // hf 0 -> dV_j , hf 1 -> dK_j (scaled)
        store_acc(hf == 0 ? T_DV : T_DK, hf == 0 ? 1.f : p.scale, dq_base + (long long)n * 3 * HD + (hf == 0 ? 2 : 1) * HD, n < N);
        if (last) {   // hf 0 -> dQ rows 0..127 , hf 1 -> dQ rows 128..255
          const int nq = hf * 128 + r;
          if (hf * 128 + q * 32 < N) store_acc(T_DQ + 64 * hf, p.scale, dq_base + (long long)nq * 3 * HD, nq < N);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_accfree);
        ++gr;
      };
      // ---- per-query vectors of this item: -lse*log2(e) and -delta = -<dO_q, O_q>
      asm volatile("bar.sync 1, 256;" ::: "memory");       // every warp is done with the previous item's vectors
      {
        const int t = threadIdx.x;                          // 0..255: one query per thread
        float nl = 0.f, nd = 0.f;
        if (t < N) {
          const bf16* po = p.o + ((long long)b * N + t) * HD + (long long)h * B2_DH;
          const bf16* pd = p.dout + ((long long)b * N + t) * HD + (long long)h * B2_DH;
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float a[8], c[8];
            V8<bf16>::load(po + 8 * k, a);
            V8<bf16>::load(pd + 8 * k, c);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc = fmaf(a[e], c[e], acc);
          }
          nd = -acc;
          nl = -p.lse[((long long)b * H + h) * N + t] * 1.4426950408889634f;
        }
        nlse_s[t] = nl;
        ndel_s[t] = nd;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int lb = 0; lb < NB; ++lb) {
        const int j = lb / NQB, i = lb % NQB;
        const int g = gb + lb, u = g & 1;
        const int nch = nq_of(i) / 16;                      // 16-column chunks in this block (1..4)
        const int c_beg = hf == 0 ? 0 : (nch + 1) / 2, c_end = hf == 0 ? (nch + 1) / 2 : nch;
        mbar_wait(bar_s0 + 8 * u, (g >> 1) & 1, 21);
        tc_fence_after();
        const uint32_t tb = T + (uint32_t)u * 128u + lane_addr;
        uint8_t* ds_row = dS_gen + ((gp & 1) * 2 + (i & 1)) * B2_CHUNK + r * 128;
        // warps whose 32 key rows are all padding skip the arithmetic: their stale P^T / dS^T rows only reach
        // accumulator rows that are never stored, or meet zero K rows in the dQ product
        const bool rows_live = j * 128 + q * 32 < N;
        for (int c = c_beg; c < c_end && rows_live; ++c) {
          uint32_t s[16], d[16];
          tmem_ld_32x16(tb + c * 16, s);
          tmem_ld_32x16(tb + 64 + c * 16, d);
          tmem_wait_ld();
          const float4* nl4 = reinterpret_cast<const float4*>(nlse_s + i * 64 + c * 16);
          const float4* nd4 = reinterpret_cast<const float4*>(ndel_s + i * 64 + c * 16);
          uint32_t pk[8], dk[8];
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const float4 l = nl4[k4], dl = nd4[k4];
            {
              float x0, x1;
              f2_unpack(f2_fma(f2_pack(__uint_as_float(s[4 * k4]), __uint_as_float(s[4 * k4 + 1])), c2, f2_pack(l.x, l.y)), x0, x1);
              const float p0 = b2_ex2(x0), p1 = b2_ex2(x1);
              float t0, t1;
              f2_unpack(f2_mul(f2_pack(p0, p1), f2_add(f2_pack(__uint_as_float(d[4 * k4]), __uint_as_float(d[4 * k4 + 1])), f2_pack(dl.x, dl.y))), t0, t1);
              pk[2 * k4] = pack_bf16(p0, p1);
              dk[2 * k4] = pack_bf16(t0, t1);
            }
            {
              float x0, x1;
              f2_unpack(f2_fma(f2_pack(__uint_as_float(s[4 * k4 + 2]), __uint_as_float(s[4 * k4 + 3])), c2, f2_pack(l.z, l.w)), x0, x1);
              const float p0 = b2_ex2(x0), p1 = b2_ex2(x1);
              float t0, t1;
              f2_unpack(f2_mul(f2_pack(p0, p1), f2_add(f2_pack(__uint_as_float(d[4 * k4 + 2]), __uint_as_float(d[4 * k4 + 3])), f2_pack(dl.z, dl.w))), t0, t1);
              pk[2 * k4 + 1] = pack_bf16(p0, p1);
              dk[2 * k4 + 1] = pack_bf16(t0, t1);
            }
          }
          tmem_st_32x8(tb + c * 16, pk);                    // P^T chunk c (bf16 pairs) over the first half of ITS OWN S^T chunk:
                                                            // the other column-half warp never reads these columns
          *reinterpret_cast<uint4*>(ds_row + (((2 * c) ^ (r & 7)) << 4)) = make_uint4(dk[0], dk[1], dk[2], dk[3]);
          *reinterpret_cast<uint4*>(ds_row + (((2 * c + 1) ^ (r & 7)) << 4)) = make_uint4(dk[4], dk[5], dk[6], dk[7]);
        }
        tmem_wait_st();
        fence_async_smem();       // generic-proxy smem writes -> visible to the UMMA operand reads
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p0 + 8 * u);
        if ((i & 1) || i == NQB - 1) ++gp;
        // the previous row's dV / dK: stored after this row's first block, so the wait for its products is covered
        if (i == 0 && j > 0) epilogue_row(j - 1, false);
      }
      epilogue_row(KT - 1, true);
      gb += NB;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == CTRL) tmem_dealloc(T, 512);
}

int attn_bwd_tc2(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                 int B, int N, int H, int dh, float scale, cudaStream_t st) {
  NRV_REQUIRE(attn_tc_supported(N, dh, NRV_BF16), "tcgen05 attention: unsupported shape N=%d dh=%d", N, dh);
  NRV_REQUIRE(((uintptr_t)qkv % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)dout % 16) == 0 && ((uintptr_t)dqkv % 16) == 0,
              "tcgen05 attention: 16-byte alignment");
  Bwd2Params p{};
  p.B = B; p.N = N; p.H = H; p.NP = (N + 15) / 16 * 16;
  p.KT = (N + 127) / 128; p.NQB = (p.NP + 63) / 64; p.items = B * H;
  p.scale = scale; p.scale_log2e = scale * 1.4426950408889634f;
  p.o = (const bf16*)out; p.dout = (const bf16*)dout; p.lse = lse; p.dqkv = (bf16*)dqkv;
  const uint64_t row_qkv = (uint64_t)3 * H * B2_DH, row_o = (uint64_t)H * B2_DH;
  CUtensorMap tq, td;
  int rc = encode_tmap_3d(&tq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, qkv, row_qkv, N, B, row_qkv * 2, row_qkv * 2 * N,
                          64, p.NP, 1, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = encode_tmap_3d(&td, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dout, row_o, N, B, row_o * 2, row_o * 2 * N,
                      64, p.NP, 1, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  const int smem = Bwd2Smem::total(p.KT, p.NP);
  NRV_REQUIRE(smem <= 227 * 1024, "tcgen05 attention backward: %d bytes of shared memory needed (N=%d)", smem, N);
  NRV_CUDA(cudaFuncSetAttribute(attn_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = p.items < num_sms() ? p.items : num_sms();
  attn_bwd2_kernel<<<grid, B2_THREADS, smem, st>>>(tq, td, p);
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

}  // namespace nrv
