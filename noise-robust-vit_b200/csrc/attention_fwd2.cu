// tcgen05 attention forward, second generation: two independent pipelines ("groups") per SM.
//   dots = q k^T * scale ; attn = softmax(dots) ; out = attn v          (simple_vit.py:70-75)
//
// The first-generation kernel (removed in round 2) walked its tiles through one serial chain
//   TMA -> S = Q K^T -> softmax -> P to smem -> O = P V -> epilogue
// with every warp in the same phase at the same time, so the MUFU pipe (exp2) idles during the MMAs and
// the tensor pipe idles during the softmax.  Here one CTA per SM runs TWO such chains on different
// (batch, head) items: each group owns 4 softmax warps (thread = score row), one control warp (TMA + MMA
// issue), half of tensor memory and its own shared-memory tiles, so one group's MMA / TMA latency is
// covered by the other group's arithmetic.  Further differences:
//   * P never touches shared memory: the softmax threads write bf16 probabilities back into tensor memory
//     (tcgen05.st, over the score columns they were computed from) and O = P V takes its A operand from
//     TMEM (tcgen05.mma, "TS" form).  No 64 KB P tile, no generic->async proxy fence on the hot path.
//   * thread = row over all NP columns (no cross-warp max / sum exchange); columns 0..127 stay in registers
//     between the max and the exp pass, the rest is read from TMEM twice
//   * the exp phases of the two groups are forced to alternate (turn barriers), so one group's MUFU burst
//     runs under the other group's MMA / TMEM / epilogue latencies
//   * exp2 arguments and row sums use packed f32x2 arithmetic; the row max uses 3-input max
//   * K of the next item is prefetched (double buffer), V is reloaded as soon as the last P V retires
//   * O leaves through swizzled staging and a TMA store (rows beyond N are clipped by the tensor map)
// TMEM per group (256 columns): S fp32 [0, NP) ; P bf16 pairs [0, NP/2) ; O fp32 [128, 192).
#include "common.cuh"
#include "nrvit_internal.h"

namespace nrv {

constexpr int F2_DH = 64;
constexpr int F2_THREADS = 32 * 10;        // warps 0-3 / 4-7: softmax of group 0 / 1 ; warps 8 / 9: control
constexpr int F2_QBYTES = 128 * 128;       // 128 rows x 64 bf16
constexpr int F2_STG = 4096;               // 32 rows x 128 B per softmax warp

struct Fwd2Params {
  int B, N, H, NP, tiles, items;
  float scale, scale_log2e;
  float* lse;   // [B, H, N] or null
  long long* dbg;   // optional phase timestamps of CTA 0: [group][tile < 32][16] clock64 values
};

struct Fwd2Smem {
  __host__ __device__ static int kv_bytes(int NP) { return (NP * 128 + 1023) & ~1023; }
  __host__ __device__ static int group_bytes(int NP) { return F2_QBYTES + 3 * kv_bytes(NP) + 4 * F2_STG; }
  __host__ __device__ static int bar_off(int NP) { return 2 * group_bytes(NP); }
  __host__ __device__ static int total(int NP) { return bar_off(NP) + 2 * 8 * 8 + 2 * 8 + 16 + 1024; }   // + 2 turn barriers
};

// (ex2f, f2_max16 and f2_exp16 live in common.cuh: shared with attention_fwd_big.cu)

// (the order-pinned primitives pv_* live in common.cuh)

template <int NCH>
__global__ void __launch_bounds__(F2_THREADS, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                 const __grid_constant__ CUtensorMap tm_out, const Fwd2Params p) {
  pdl_launch_dependents();   // the next kernel in the stream (a PDL-launched GEMM) may start its prologue
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const int NP = p.NP, N = p.N, H = p.H;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool ctrl = warp >= 8;
  const int g = ctrl ? (warp - 8) & 1 : warp >> 2;      // pipeline group of this warp
  const int KV = Fwd2Smem::kv_bytes(NP);
  const uint32_t sG = sbase + g * Fwd2Smem::group_bytes(NP);
  const uint32_t sQ = sG, sK0 = sG + F2_QBYTES, sV = sK0 + 2 * KV;
  const uint32_t bar0 = sbase + Fwd2Smem::bar_off(NP) + g * 64;
  const uint32_t bar_q = bar0, bar_k0 = bar0 + 8, bar_v = bar0 + 24, bar_s = bar0 + 32, bar_p = bar0 + 40,
                 bar_o = bar0 + 48, bar_free = bar0 + 56;   // bar_k1 = bar_k0 + 8
  const uint32_t bar_turn0 = sbase + Fwd2Smem::bar_off(NP) + 128;   // [g]: group g may start its exp phase
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + Fwd2Smem::bar_off(NP) + 144);

  if (warp == 8) {
    if (elect_one()) {
      tma_prefetch_desc(&tm_q);
      tma_prefetch_desc(&tm_kv);
      tma_prefetch_desc(&tm_out);
      for (int gg = 0; gg < 2; ++gg) {
        const uint32_t b = sbase + Fwd2Smem::bar_off(NP) + gg * 64;
        mbar_init(b, 1);            // q
        mbar_init(b + 8, 1);        // k0
        mbar_init(b + 16, 1);       // k1
        mbar_init(b + 24, 1);       // v
        mbar_init(b + 32, 1);       // s
        mbar_init(b + 40, 4);       // p
        mbar_init(b + 48, 1);       // o
        mbar_init(b + 56, 4);       // free
        mbar_init(bar_turn0 + 8 * gg, 4);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t T = *tmem_ptr_smem + (uint32_t)g * 256u;   // this group's half of tensor memory

  // items of this group: vc, vc + nvc, ...  (group 1 continues where group 0's first round ends, so an SM gets
  // ceil or floor of items / SMs, not twice the remainder)
  const int vc = g * (int)gridDim.x + (int)blockIdx.x, nvc = 2 * (int)gridDim.x;
  const int my_items = vc < p.items ? (p.items - vc + nvc - 1) / nvc : 0;
  const int total_tiles = my_items * p.tiles;
  const int ksteps = NP / 16;
  const int vc_o = (1 - g) * (int)gridDim.x + (int)blockIdx.x;   // the other group of this CTA
  const int partner_tiles = (vc_o < p.items ? (p.items - vc_o + nvc - 1) / nvc : 0) * p.tiles;

  if (ctrl) {
    // ================================ TMA + MMA issue (one lane per group) ======================
    if (elect_one() && total_tiles > 0) {
      const uint32_t idesc_s = make_idesc(1u, 0u, 0u, 128u, (uint32_t)NP);
      const uint32_t idesc_o = make_idesc(1u, 0u, 1u, 128u, 64u);
      // (batch, head) of items li and li + 1, kept in a two-entry table: one pair of integer divisions per ITEM on this
      // thread instead of four per tile (ncu source view of the backward kernel: ~100 cycles of dependent latency each)
      int cb0 = 0, ch0 = 0, cb1 = 0, ch1 = 0;             // (scalars, not an indexed array: that would live in local memory)
      auto fill = [&](int li) {
        const int item = vc + li * nvc;
        const int b = item / H, h = item - b * H;
        if (li & 1) { cb1 = b; ch1 = h; } else { cb0 = b; ch0 = h; }
      };
      auto coords = [&](int li, int& b, int& h) { b = (li & 1) ? cb1 : cb0; h = (li & 1) ? ch1 : ch0; };
      fill(0);
      auto issue_q = [&](int li_n, int t_n) {
        int b, h; coords(li_n, b, h);
        mbar_arrive_expect_tx(bar_q, F2_QBYTES);
        tma_load_3d(sQ, &tm_q, bar_q, (0 * H + h) * F2_DH, t_n * 128, b);
      };
      auto issue_k = [&](int li) {
        int b, h; coords(li, b, h);
        const uint32_t bar = bar_k0 + 8 * (li & 1);
        mbar_arrive_expect_tx(bar, NP * 128);
        tma_load_3d(sK0 + (li & 1) * KV, &tm_kv, bar, (1 * H + h) * F2_DH, 0, b);
      };
      auto issue_v = [&](int li) {
        int b, h; coords(li, b, h);
        mbar_arrive_expect_tx(bar_v, NP * 128);
        tma_load_3d(sV, &tm_kv, bar_v, (2 * H + h) * F2_DH, 0, b);
      };
      issue_q(0, 0); issue_k(0); issue_v(0);
      for (int gi = 0, li = 0, t = 0; gi < total_tiles; ++gi, (++t == p.tiles ? (t = 0, ++li) : 0)) {
        if (t == 0 && li + 1 < my_items) fill(li + 1);
        const uint32_t ph = gi & 1;
        long long* dbg = (p.dbg != nullptr && blockIdx.x == 0 && gi < 32) ? p.dbg + (g * 32 + gi) * 16 : nullptr;
        if (dbg) dbg[0] = clock64();
        // K of the next item: its buffer was last read by the S MMAs of item li - 1, all retired (bar_s waited)
        if (t == 0 && li + 1 < my_items) issue_k(li + 1);
        mbar_wait(bar_q, ph, 10);
        if (t == 0) mbar_wait(bar_k0 + 8 * (li & 1), (li >> 1) & 1, 11);
        if (gi > 0) mbar_wait(bar_free, (gi - 1) & 1, 14);   // previous tile's O drained: the TMEM half is free
        tc_fence_after();
        if (dbg) dbg[1] = clock64();
        {
          const uint64_t a1 = make_smem_desc_sw128(sQ, 16, 1024);
          const uint64_t b1 = make_smem_desc_sw128(sK0 + (li & 1) * KV, 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(T, a1 + 2 * k, b1 + 2 * k, idesc_s, k > 0);
          umma_commit(bar_s);
        }
        mbar_wait(bar_s, ph, 12);               // S done: the Q tile may be overwritten
        if (dbg) dbg[2] = clock64();
        if (gi + 1 < total_tiles) issue_q(t + 1 == p.tiles ? li + 1 : li, t + 1 == p.tiles ? 0 : t + 1);
        mbar_wait(bar_p, ph, 13);               // P written to TMEM by the softmax warps
        if (t == 0) mbar_wait(bar_v, li & 1, 15);
        tc_fence_after();
        if (dbg) dbg[3] = clock64();
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t bd = make_smem_desc_sw128(sV + ks * 2048, (uint32_t)(NP * 128), 1024);
          umma_bf16_ts(T + 128, T + ks * 8, bd, idesc_o, ks > 0);
        }
        umma_commit(bar_o);
        if (dbg) dbg[4] = clock64();
        if (t == p.tiles - 1 && li + 1 < my_items) {
          mbar_wait(bar_o, ph, 16);             // last P V of this item retired: V may be overwritten
          issue_v(li + 1);
        }
      }
    }
    __syncwarp();
  } else {
    // ================================ softmax + epilogue warps =================================
    const int q = warp & 3;                               // TMEM lane quarter
    const int r = q * 32 + lane;                          // row within the 128-row tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t T_S = T + lane_addr, T_O = T + 128 + lane_addr;
    uint8_t* stg = smem + g * Fwd2Smem::group_bytes(NP) + F2_QBYTES + 3 * KV + q * F2_STG;
    constexpr int nch = NCH;
    const uint64_t c2 = f2_pack(p.scale_log2e, p.scale_log2e);
    const uint32_t bar_turn_mine = bar_turn0 + 8 * g, bar_turn_other = bar_turn0 + 8 * (1 - g);
    int b = 0, h = 0;
    for (int gi = 0, li = 0, t = 0; gi < total_tiles; ++gi, (++t == p.tiles ? (t = 0, ++li) : 0)) {
      if (t == 0) {                                       // one division pair per item, not four per tile
        const int item = vc + li * nvc;
        b = item / H; h = item % H;
      }
      const uint32_t ph = gi & 1;
      const int n = t * 128 + r;                          // token index of this thread's row
      const bool warp_active = t * 128 + q * 32 < N;      // warp-uniform: any valid row in this warp
      long long* sdbg = (p.dbg != nullptr && blockIdx.x == 0 && gi < 32 && q == 0 && lane == 0) ? p.dbg + (g * 32 + gi) * 16 + 8 : nullptr;
      if (sdbg) sdbg[0] = clock64();
      mbar_wait_inline(bar_s, ph);   // (call-free waits: the score row lives in registers)
      tc_fence_after();
      if (sdbg) sdbg[1] = clock64();
      float mx = 0.f, tot = 1.f;
      // The exp phases of the two groups alternate (g0 tile k, g1 tile k, g0 tile k+1, ...): while one group
      // saturates the MUFU pipe, the other sits in its MMA / TMEM / epilogue latencies.
      const bool turn_wait = g == 0 ? (gi >= 1 && gi - 1 < partner_tiles) : (gi < partner_tiles);
      const uint32_t turn_ph = g == 0 ? (gi - 1) & 1 : gi & 1;
      if (warp_active) {
        // Scores: 16-column chunks 0..7 (columns 0..127) are read once and stay in registers for both passes;
        // chunks 8..12 (columns 128..207) are read for the max and again for the exp (10 warps -> 168 registers
        // per thread: the whole row does not fit).  Three TMEM round trips, the third under the exp of the
        // held chunks.
        uint32_t held[8][16], tail[5][16];
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int c = 8; c < 13; ++c)
          if (c < nch) tmem_ld_32x16(T_S + c * 16, tail[c - 8]);
#pragma unroll
        for (int c = 0; c < 2; ++c)
          if (c < nch) tmem_ld_32x16(T_S + c * 16, held[c]);
        tmem_wait_ld();
#pragma unroll
        for (int c = 8; c < 13; ++c)
          if (c < nch) f2_max16(tail[c - 8], c * 16, N, m0, m1);
#pragma unroll
        for (int c = 2; c < 8; ++c)
          if (c < nch) tmem_ld_32x16(T_S + c * 16, held[c]);
#pragma unroll
        for (int c = 0; c < 2; ++c)
          if (c < nch) f2_max16(held[c], c * 16, N, m0, m1);
        tmem_wait_ld();
#pragma unroll
        for (int c = 2; c < 8; ++c)
          if (c < nch) f2_max16(held[c], c * 16, N, m0, m1);
        mx = fmaxf(m0, m1);
        // p = exp2(s * c - mx * c) ; row sum ; bf16 pairs back into TMEM over the consumed scores (P occupies
        // columns [0, NP/2) <= 104, never the tail chunks at columns >= 128 that are re-read below)
        if (sdbg) sdbg[2] = clock64();
        if (turn_wait) mbar_wait_inline(bar_turn_mine, turn_ph);
        const float noff = -mx * p.scale_log2e;
        const uint64_t noff2 = f2_pack(noff, noff);
        uint64_t sum2 = f2_pack(0.f, 0.f), sum2b = f2_pack(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 16; j += 2) pv_fma2(held[0][j], held[0][j + 1], c2, noff2);
#pragma unroll
        for (int k = 0; k <= nch; ++k) {
          if (k == 5 && nch > 8) {                         // chunks 0..3 are consumed: their registers take the re-read
#pragma unroll
            for (int c = 8; c < 13; ++c)
              if (c < nch) tmem_ld_32x16(T_S + c * 16, tail[c - 8]);
          }
          if (k + 1 == 8 && nch > 8) tmem_wait_ld();       // stage A reaches the re-read chunks
          uint32_t (&vn)[16] = (k + 1 < 8) ? held[(k + 1 < 8) ? k + 1 : 0] : tail[(k + 1 >= 8 && k + 1 < 13) ? k + 1 - 8 : 0];
          uint32_t (&vc)[16] = (k < 8) ? held[(k < 8) ? k : 0] : tail[(k >= 8 && k < 13) ? k - 8 : 0];
          uint32_t (&vp)[16] = (k - 1 < 8) ? held[(k >= 1 && k - 1 < 8) ? k - 1 : 0] : tail[(k - 1 >= 8 && k - 1 < 13) ? k - 1 - 8 : 0];
          if (k >= 1 && k == nch && (N & 15) != 0) {        // padding columns of the row's last chunk
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if ((k - 1) * 16 + j >= N) vp[j] = 0u;
          }
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (k < nch) { pv_ex2(vc[2 * j]); pv_ex2(vc[2 * j + 1]); }
            if (k >= 1) {
              if (j & 1) pv_add2(sum2b, vp[2 * j], vp[2 * j + 1]); else pv_add2(sum2, vp[2 * j], vp[2 * j + 1]);
              pk[j] = pv_pack(vp[2 * j], vp[2 * j + 1]);
            }
            if (k + 1 < nch) pv_fma2(vn[2 * j], vn[2 * j + 1], c2, noff2);
          }
          if (k >= 1) tmem_st_32x8(T_S + (k - 1) * 8, pk);
        }
        sum2 = f2_add(sum2, sum2b);
        float s0, s1;
        f2_unpack(sum2, s0, s1);
        tot = s0 + s1;
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_turn_other);
        tmem_wait_st();
      } else {
        if (turn_wait) mbar_wait_inline(bar_turn_mine, turn_ph);
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_turn_other);
      }
      if (sdbg) sdbg[3] = clock64();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
      // off the critical path: normaliser and log-sum-exp while P V runs
      const float inv = __fdividef(1.f, tot);
      if (warp_active && n < N && p.lse) p.lse[((long long)b * H + h) * N + n] = mx * p.scale + __logf(tot);

      if (sdbg) sdbg[4] = clock64();
      mbar_wait_inline(bar_o, ph);
      tc_fence_after();
      if (sdbg) sdbg[5] = clock64();
      if (!warp_active) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_free);
      } else {
        uint32_t vo[2][32];
        tmem_ld_32x32(T_O, vo[0]);
        tmem_ld_32x32(T_O + 32, vo[1]);
        tmem_wait_ld();
        // O is in registers: hand the TMEM half back before the arithmetic and the store
        tc_fence_before();
        __syncwarp();
        if (sdbg) sdbg[6] = clock64();
        if (lane == 0) mbar_arrive(bar_free);
        // ---- epilogue: O / rowsum -> bf16 -> swizzled staging -> TMA store (32 rows x 64 columns per warp)
        if (lane == 0) tma_store_wait_read<0>();      // the previous store has finished reading the staging
        __syncwarp();
        const uint64_t inv2 = f2_pack(inv, inv);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const uint32_t (&v)[32] = vo[hh];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float a, c;
              f2_unpack(f2_mul(f2_pack(__uint_as_float(v[8 * u + 2 * j]), __uint_as_float(v[8 * u + 2 * j + 1])), inv2), a, c);
              w[j] = pack_bf16(a, c);
            }
            sts128(smem_u32(stg) + lane * 128 + (((hh * 4 + u) ^ (lane & 7)) << 4), w[0], w[1], w[2], w[3]);
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tm_out, smem_u32(stg), h * F2_DH, t * 128 + q * 32, b);
          tma_store_commit();
        }
      }
    }
    if (lane == 0) tma_store_wait<0>();   // smem must outlive the last bulk store
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(*tmem_ptr_smem, 512);
}

int attn_fwd_tc2(const void* qkv, void* out, float* lse, int B, int N, int H, int dh, float scale, cudaStream_t st) {
  NRV_REQUIRE(attn_tc_supported(N, dh, NRV_BF16), "tcgen05 attention: unsupported shape N=%d dh=%d", N, dh);
  NRV_REQUIRE(((uintptr_t)qkv % 16) == 0 && ((uintptr_t)out % 16) == 0, "tcgen05 attention: 16-byte alignment");
  Fwd2Params p{};
  p.B = B; p.N = N; p.H = H; p.NP = (N + 15) / 16 * 16;
  p.tiles = (N + 127) / 128; p.items = B * H;
  p.scale = scale; p.scale_log2e = scale * 1.4426950408889634f;
  p.lse = lse;
  p.dbg = attn_tc_get_debug();
  const uint64_t row_qkv = (uint64_t)3 * H * F2_DH, row_o = (uint64_t)H * F2_DH;
  CUtensorMap tq, tkv, to;
  int rc = encode_tmap_3d(&tq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, qkv, row_qkv, N, B, row_qkv * 2, row_qkv * 2 * N,
                          64, 128, 1, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = encode_tmap_3d(&tkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, qkv, row_qkv, N, B, row_qkv * 2, row_qkv * 2 * N,
                      64, p.NP, 1, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = encode_tmap_3d(&to, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, out, row_o, N, B, row_o * 2, row_o * 2 * N,
                      64, 32, 1, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  const int smem = Fwd2Smem::total(p.NP);
  NRV_REQUIRE(smem <= 227 * 1024, "tcgen05 attention: %d bytes of shared memory needed (N=%d)", smem, N);
  const int units = (p.items + 1) / 2;
  const int grid = units < num_sms() ? units : num_sms();
  switch (p.NP / 16) {
#define F2_CASE(n)                                                                                                  \
    case n:                                                                                                         \
      NRV_CUDA(cudaFuncSetAttribute(attn_fwd2_kernel<n>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));       \
      attn_fwd2_kernel<n><<<grid, F2_THREADS, smem, st>>>(tq, tkv, to, p);                                          \
      break;
    F2_CASE(1) F2_CASE(2) F2_CASE(3) F2_CASE(4) F2_CASE(5) F2_CASE(6) F2_CASE(7) F2_CASE(8) F2_CASE(9) F2_CASE(10)
    F2_CASE(11) F2_CASE(12) F2_CASE(13)
#undef F2_CASE
    default:
      NRV_REQUIRE(false, "tcgen05 attention: N=%d out of range", N);
  }
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

}  // namespace nrv
