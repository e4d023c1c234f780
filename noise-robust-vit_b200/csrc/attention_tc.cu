// tcgen05 attention for short sequences (N <= 256 tokens, head dim 64): forward, and backward as two
// recompute passes (dQ pass, dK/dV pass).  Production bf16 path of
//   dots = q k^T * scale ; attn = softmax(dots) ; out = attn v          (simple_vit.py:70-75)
// and of its autograd.
//
// One persistent CTA per SM walks (batch, head) items.  Because N <= 256 the whole score row of a
// 128-row tile fits in TMEM (<= 256 fp32 columns), so the softmax is a plain two-pass row softmax
// with one thread per row (tcgen05.ld 32x32b: lane == row, no shuffles), not an online softmax:
//
//   pass      row tile (A)      column tensors (B)     MMA 1/2 (K = dh)            elementwise -> smem (bf16)     output MMAs (K = N)
//   FWD       Q_t               K, V                   S = Q_t K^T                 P = exp2(S - max)              O  = P V
//   DQ        Q_t, dO_t         K, V                   S = Q_t K^T, dP = dO_t V^T  dS = P o (dP - delta) * scale  dQ = dS K
//   DKV       K_j, V_j          Q, dO                  S^T = K_j Q^T, dP^T = V_j dO^T   P^T ; dS^T                dV = P^T dO ; dK = dS^T Q
//
// Operands arrive by TMA (3-D maps over the packed [B, N, 3, H, dh] projection output, so rows
// beyond N are zero-filled per image), sit in 128B-swizzled smem, and feed tcgen05.mma directly:
// K-major for MMA 1/2, and the same column tiles re-read MN-major for the output MMAs (V, K, dO, Q
// are [tokens, dh] row-major = [K, N]).  P / dS tiles are written by the softmax threads in the
// canonical K-major SW128 layout.  Accumulators live in TMEM; outputs overlay the score columns
// once the softmax threads are done with them.
// Warps 0-7: softmax + epilogue; warp w works on TMEM lanes 32(w%4).. and on column half w/4, so two
// warps share every row (row max / row sum are exchanged through shared memory).  Warp 8: TMA + MMA
// issue (one lane).
#include "common.cuh"
#include "nrvit_internal.h"

#include <stdlib.h>

namespace nrv {

enum { ATT_FWD = 0, ATT_DQ = 1, ATT_DKV = 2 };

constexpr int ATT_SM_WARPS = 8;                 // softmax / epilogue warps
constexpr int ATT_THREADS = 32 * (ATT_SM_WARPS + 1);
constexpr int ATT_MAXCH = 7;                    // 16-column chunks per thread (column half <= 112)
constexpr int ATT_DH = 64;
constexpr int ROW_TILE_BYTES = 128 * 128;  // 128 rows x 64 bf16

struct AttnParams {
  int B, N, H, NP;      // NP = N rounded up to 16 (<= 256)
  int tiles;            // 128-row tiles per head
  int items;            // B * H
  float scale;          // dh^-0.5
  float scale_log2e;    // scale * log2(e)
  const bf16* o;        // [B, N, H*dh]   (DQ: delta = rowsum(dO o O))
  const bf16* dout;     // [B, N, H*dh]
  bf16* out;            // FWD
  bf16* dqkv;           // DQ / DKV
  float* lse;           // [B, H, N]  FWD writes, bwd reads
  float* delta;         // [B, H, N]  DQ writes, DKV reads
  long long* dbg;       // optional: control-thread phase timestamps of CTA 0 ([tile][8] clock64 values)
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 16 consecutive bf16 of row r starting at column k0 (k0 % 16 == 0) of a K-major SW128 tile made
// of [128 rows x 64 cols] chunks
__device__ __forceinline__ void store_p16(uint8_t* tile, int r, int k0, const float (&v)[16]) {
  uint8_t* row = tile + (k0 >> 6) * ROW_TILE_BYTES + r * 128;
  const int u = (k0 & 63) >> 3;
  *reinterpret_cast<uint4*>(row + (((u) ^ (r & 7)) << 4)) =
      make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  *reinterpret_cast<uint4*>(row + (((u + 1) ^ (r & 7)) << 4)) =
      make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
}

template <int MODE>
struct AttSmem {
  static constexpr int NROW = MODE == ATT_FWD ? 1 : 2;   // row tiles
  static constexpr int NP_ = MODE == ATT_DKV ? 2 : 1;    // P-like tiles
  static constexpr int NCOLBUF = MODE == ATT_FWD ? 2 : 1;  // FWD double-buffers K/V across (b, h) items
  __host__ __device__ static int col_bytes(int NP) { return (NP * 128 + 1023) & ~1023; }
  __host__ __device__ static int p_bytes(int NP) { return ((NP + 63) / 64) * ROW_TILE_BYTES; }
  __host__ __device__ static int off_col(int NP, int i) { return NROW * ROW_TILE_BYTES + i * col_bytes(NP); }
  __host__ __device__ static int off_p(int NP, int i) { return off_col(NP, 2 * NCOLBUF) + i * p_bytes(NP); }
  __host__ __device__ static int off_vec(int NP) { return off_p(NP, NP_); }
  __host__ __device__ static int off_xch(int NP) { return off_vec(NP) + (MODE == ATT_DKV ? 2 * 256 * 4 : 0); }
  __host__ __device__ static int off_bar(int NP) { return off_xch(NP) + 4 * 128 * 4; }
  __host__ __device__ static int total(int NP) { return off_bar(NP) + 8 * 8 + 16 + 1024; }   // 8 barriers + tmem ptr
};

template <int MODE>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv_row, const __grid_constant__ CUtensorMap tm_qkv_col,
               const __grid_constant__ CUtensorMap tm_do_row, const __grid_constant__ CUtensorMap tm_do_col,
               const AttnParams p) {
  using L = AttSmem<MODE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const int NP = p.NP;
  const uint32_t sR1 = sbase, sR2 = sbase + ROW_TILE_BYTES;
  const uint32_t sC1 = sbase + L::off_col(NP, 0), sC2 = sbase + L::off_col(NP, 1);
  uint8_t* P1 = smem + L::off_p(NP, 0);
  uint8_t* P2 = smem + L::off_p(NP, 1);
  const uint32_t sP1 = sbase + L::off_p(NP, 0), sP2 = sbase + L::off_p(NP, 1);
  float* lse_s = reinterpret_cast<float*>(smem + L::off_vec(NP));
  float* del_s = lse_s + 256;
  const uint32_t bar0 = sbase + L::off_bar(NP);
  const uint32_t bar_c = bar0, bar_r = bar0 + 8, bar_s = bar0 + 16, bar_p = bar0 + 24, bar_o = bar0 + 32,
                 bar_free = bar0 + 40, bar_c2 = bar0 + 48;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + L::off_bar(NP) + 64);

  float* max_s = reinterpret_cast<float*>(smem + L::off_xch(NP));   // [2][128]
  float* sum_s = max_s + 256;                                       // [2][128]
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int CTRL = ATT_SM_WARPS;

  if (warp == CTRL) {
    if (elect_one()) {
      tma_prefetch_desc(&tm_qkv_row);
      tma_prefetch_desc(&tm_qkv_col);
      if (MODE != ATT_FWD) { tma_prefetch_desc(&tm_do_row); tma_prefetch_desc(&tm_do_col); }
      mbar_init(bar_c, 1);
      mbar_init(bar_c2, 1);
      mbar_init(bar_r, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_p, ATT_SM_WARPS);
      mbar_init(bar_o, 1);
      mbar_init(bar_free, ATT_SM_WARPS);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_smem;
  const uint32_t T_S = tmem, T_S2 = tmem + 256;
  const uint32_t T_OUT1 = MODE == ATT_FWD ? tmem + 256 : tmem;   // O | dQ | dV
  const uint32_t T_OUT2 = tmem + 64;                             // dK (DKV)

  const int H = p.H, N = p.N;
  const int my_items = (p.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total_tiles = my_items * p.tiles;
  const int ksteps = NP / 16;

  if (warp == CTRL) {
    // ================================ TMA + MMA issue (one lane) ================================
    if (elect_one()) {
      const uint32_t idesc_s = make_idesc(1u, 0u, 0u, 128u, (uint32_t)NP);
      const uint32_t idesc_o = make_idesc(1u, 0u, 1u, 128u, 64u);
      auto issue_rows = [&](int g) {
        const int item = blockIdx.x + (g / p.tiles) * gridDim.x;
        const int t = g % p.tiles;
        const int b = item / H, h = item % H;
        mbar_arrive_expect_tx(bar_r, L::NROW * ROW_TILE_BYTES);
        if (MODE == ATT_FWD) {
          tma_load_3d(sR1, &tm_qkv_row, bar_r, (0 * H + h) * ATT_DH, t * 128, b);
        } else if (MODE == ATT_DQ) {
          tma_load_3d(sR1, &tm_qkv_row, bar_r, (0 * H + h) * ATT_DH, t * 128, b);
          tma_load_3d(sR2, &tm_do_row, bar_r, h * ATT_DH, t * 128, b);
        } else {
          tma_load_3d(sR1, &tm_qkv_row, bar_r, (1 * H + h) * ATT_DH, t * 128, b);
          tma_load_3d(sR2, &tm_qkv_row, bar_r, (2 * H + h) * ATT_DH, t * 128, b);
        }
      };
      auto issue_cols = [&](int item) {
        const int b = item / H, h = item % H;
        mbar_arrive_expect_tx(bar_c, 2 * NP * 128);
        if (MODE == ATT_DKV) {
          tma_load_3d(sC1, &tm_qkv_col, bar_c, (0 * H + h) * ATT_DH, 0, b);   // Q
          tma_load_3d(sC2, &tm_do_col, bar_c, h * ATT_DH, 0, b);              // dO
        } else {
          tma_load_3d(sC1, &tm_qkv_col, bar_c, (1 * H + h) * ATT_DH, 0, b);   // K
          tma_load_3d(sC2, &tm_qkv_col, bar_c, (2 * H + h) * ATT_DH, 0, b);   // V
        }
      };
      if (MODE == ATT_FWD) {
        // Forward: the control thread runs ahead of the softmax warps.  S is consumed into registers before
        // bar_p, O lives in its own TMEM columns, and K/V are double buffered across items, so the next
        // tile's S = Q K^T is issued as soon as the current P V is, and the epilogue of tile g overlaps it.
        const int CB = L::col_bytes(NP);
        auto issue_cols_buf = [&](int li) {
          const int item = blockIdx.x + li * gridDim.x;
          const int b = item / H, h = item % H;
          const uint32_t bar = (li & 1) ? bar_c2 : bar_c;
          const uint32_t base = sC1 + (li & 1) * 2 * CB;
          mbar_arrive_expect_tx(bar, 2 * NP * 128);
          tma_load_3d(base, &tm_qkv_col, bar, (1 * H + h) * ATT_DH, 0, b);        // K
          tma_load_3d(base + CB, &tm_qkv_col, bar, (2 * H + h) * ATT_DH, 0, b);   // V
        };
        if (total_tiles > 0) { issue_rows(0); issue_cols_buf(0); }
        for (int g = 0; g < total_tiles; ++g) {
          const int li = g / p.tiles, t = g % p.tiles;
          const uint32_t ph = g & 1;
          const uint32_t cbuf = sC1 + (li & 1) * 2 * CB;
          if (t == 0 && li + 1 < my_items) {
            // the other K/V buffer was last read by the P V of the previous item's final tile
            if (g > 0) mbar_wait(bar_o, (g - 1) & 1, 15);
            issue_cols_buf(li + 1);
          }
          mbar_wait(bar_r, ph, 10);
          if (t == 0) mbar_wait((li & 1) ? bar_c2 : bar_c, (li >> 1) & 1, 11);
          tc_fence_after();
          {
            const uint64_t a1 = make_smem_desc_sw128(sR1, 16, 1024), b1 = make_smem_desc_sw128(cbuf, 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(T_S, a1 + 2 * k, b1 + 2 * k, idesc_s, k > 0);
            umma_commit(bar_s);
          }
          mbar_wait(bar_s, ph, 12);             // S done: the Q tile may be overwritten
          if (g + 1 < total_tiles) issue_rows(g + 1);
          mbar_wait(bar_p, ph, 13);             // P written (and S consumed, previous O drained)
          tc_fence_after();
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t a = make_smem_desc_sw128(sP1 + (ks >> 2) * ROW_TILE_BYTES + (ks & 3) * 32, 16, 1024);
            const uint64_t bd = make_smem_desc_sw128(cbuf + CB + ks * 2048, (uint32_t)(NP * 128), 1024);
            umma_bf16(T_OUT1, a, bd, idesc_o, ks > 0);
          }
          umma_commit(bar_o);
        }
      } else {
      if (total_tiles > 0) issue_rows(0);
      for (int g = 0; g < total_tiles; ++g) {
        const int li = g / p.tiles, t = g % p.tiles;
        const uint32_t ph = g & 1;
        long long* dbg = (p.dbg != nullptr && blockIdx.x == 0 && g < 64) ? p.dbg + g * 8 : nullptr;
        if (dbg) dbg[0] = clock64();
        if (t == 0) issue_cols(blockIdx.x + li * gridDim.x);
        mbar_wait(bar_r, ph, 10);
        if (dbg) dbg[1] = clock64();
        if (t == 0) mbar_wait(bar_c, li & 1, 11);
        if (dbg) dbg[2] = clock64();
        tc_fence_after();
        {
          const uint64_t a1 = make_smem_desc_sw128(sR1, 16, 1024), b1 = make_smem_desc_sw128(sC1, 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(T_S, a1 + 2 * k, b1 + 2 * k, idesc_s, k > 0);
          if (MODE != ATT_FWD) {
            const uint64_t a2 = make_smem_desc_sw128(sR2, 16, 1024), b2 = make_smem_desc_sw128(sC2, 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(T_S2, a2 + 2 * k, b2 + 2 * k, idesc_s, k > 0);
          }
          umma_commit(bar_s);
        }
        if (dbg) dbg[3] = clock64();
        mbar_wait(bar_s, ph, 12);             // MMA 1/2 retired: the row tiles may be overwritten
        if (dbg) dbg[4] = clock64();
        if (g + 1 < total_tiles) issue_rows(g + 1);
        mbar_wait(bar_p, ph, 13);             // P / dS tiles written by the softmax warps
        if (dbg) dbg[5] = clock64();
        tc_fence_after();
        {
          const uint32_t colB1 = MODE == ATT_FWD ? sC2 : (MODE == ATT_DQ ? sC1 : sC2);  // V | K | dO
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t a = make_smem_desc_sw128(sP1 + (ks >> 2) * ROW_TILE_BYTES + (ks & 3) * 32, 16, 1024);
            const uint64_t bd = make_smem_desc_sw128(colB1 + ks * 2048, (uint32_t)(NP * 128), 1024);
            umma_bf16(T_OUT1, a, bd, idesc_o, ks > 0);
          }
          if (MODE == ATT_DKV) {
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint64_t a = make_smem_desc_sw128(sP2 + (ks >> 2) * ROW_TILE_BYTES + (ks & 3) * 32, 16, 1024);
              const uint64_t bd = make_smem_desc_sw128(sC1 + ks * 2048, (uint32_t)(NP * 128), 1024);   // Q
              umma_bf16(T_OUT2, a, bd, idesc_o, ks > 0);
            }
          }
          umma_commit(bar_o);
        }
        if (dbg) dbg[6] = clock64();
        mbar_wait(bar_free, ph, 14);          // epilogue drained TMEM; smem tiles reusable
        if (dbg) dbg[7] = clock64();
      }
      }  // MODE != ATT_FWD
    }
    __syncwarp();
  } else {
    // ================================ softmax + epilogue warps =================================
    const int q = warp & 3;                               // TMEM lane quarter
    const int hf = warp >> 2;                             // column half
    const int r = q * 32 + lane;                          // row within the 128-row tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const long long HD = (long long)H * ATT_DH;
    // column split at a multiple of 16: [0, C0) for hf 0, [C0, NP) for hf 1
    const int C0 = ((NP + 31) / 32) * 16;
    const int cbeg = hf == 0 ? 0 : C0;
    const int nch = ((hf == 0 ? C0 : NP) - cbeg) / 16;    // <= ATT_MAXCH
    for (int g = 0; g < total_tiles; ++g) {
      const int li = g / p.tiles, t = g % p.tiles;
      const int item = blockIdx.x + li * gridDim.x;
      const int b = item / H, h = item % H;
      const uint32_t ph = g & 1;
      const int n = t * 128 + r;                          // token index of this thread's row
      const bool row_ok = n < N;
      float row_lse = 0.f, row_delta = 0.f;
      if (MODE == ATT_DQ) {
        if (row_ok) {
          row_lse = p.lse[((long long)b * H + h) * N + n] * 1.4426950408889634f;
          const bf16* po = p.o + ((long long)b * N + n) * HD + (long long)h * ATT_DH;
          const bf16* pd = p.dout + ((long long)b * N + n) * HD + (long long)h * ATT_DH;
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float a[8], c[8];
            V8<bf16>::load(po + 8 * j, a);
            V8<bf16>::load(pd + 8 * j, c);
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) acc = fmaf(a[qq], c[qq], acc);
          }
          row_delta = acc;
          if (hf == 0) p.delta[((long long)b * H + h) * N + n] = acc;
        }
      }
      if (MODE == ATT_DKV && t == 0) {
        // per-query vectors of this (b, h): safe to overwrite, every warp passed bar_o of the previous tile
        for (int i = threadIdx.x; i < NP; i += ATT_SM_WARPS * 32) {
          const bool ok = i < N;
          lse_s[i] = ok ? p.lse[((long long)b * H + h) * N + i] * 1.4426950408889634f : 0.f;
          del_s[i] = ok ? p.delta[((long long)b * H + h) * N + i] : 0.f;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      long long* sdbg = (p.dbg != nullptr && blockIdx.x == 0 && g < 64 && threadIdx.x == 0) ? p.dbg + 512 + g * 8 : nullptr;
      mbar_wait(bar_s, ph, 20);
      tc_fence_after();
      if (sdbg) sdbg[0] = clock64();

      if (MODE == ATT_FWD) {
        // ---- scores of this thread's column half stay in registers: TMEM is read once
        uint32_t sv[ATT_MAXCH][16];
#pragma unroll
        for (int c = 0; c < ATT_MAXCH; ++c)
          if (c < nch) tmem_ld_32x16(T_S + lane_addr + cbeg + c * 16, sv[c]);
        tmem_wait_ld();
        if (sdbg) sdbg[1] = clock64();
        // only the last 16-column chunk of the row can hold padding columns (NP - N < 16): every other
        // chunk runs without per-element masks
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < ATT_MAXCH; ++c)
          if (c < nch) {
            if (cbeg + c * 16 + 16 <= N) {
#pragma unroll
              for (int j = 0; j < 16; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(sv[c][j]));
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (cbeg + c * 16 + j < N) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(sv[c][j]));
            }
          }
        float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        max_s[hf * 128 + r] = mx;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        mx = fmaxf(max_s[r], max_s[128 + r]);
        if (sdbg) sdbg[2] = clock64();
        const float moff = mx * p.scale_log2e;
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < ATT_MAXCH; ++c)
          if (c < nch) {
            float e[16];
            if (cbeg + c * 16 + 16 <= N) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                e[j] = ex2(fmaf(__uint_as_float(sv[c][j]), p.scale_log2e, -moff));
                s4[j & 3] += e[j];
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float xx = ex2(fmaf(__uint_as_float(sv[c][j]), p.scale_log2e, -moff));
                e[j] = (cbeg + c * 16 + j < N) ? xx : 0.f;
                s4[j & 3] += e[j];
              }
            }
            store_p16(P1, r, cbeg + c * 16, e);
          }
        const float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
        sum_s[hf * 128 + r] = sum;
        if (hf == 0) max_s[r] = mx;   // keep the row max for the lse (both halves hold the same value)
      } else {
        // ---- backward passes: chunk loop with the next chunk's TMEM loads in flight
        uint32_t va[16], wa[16], vb[16], wb[16];
        auto process = [&](const uint32_t (&v)[16], const uint32_t (&w)[16], int c0) {
          float pe[16], ds[16];
          const bool full = c0 + 16 <= N;   // warp-uniform: only the row's last chunk holds padding columns
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float l2 = MODE == ATT_DQ ? row_lse : lse_s[c0 + j];
            const float dl = MODE == ATT_DQ ? row_delta : del_s[c0 + j];
            float xx = ex2(fmaf(__uint_as_float(v[j]), p.scale_log2e, -l2));
            if (!full && c0 + j >= N) xx = 0.f;
            pe[j] = xx;
            ds[j] = xx * ((__uint_as_float(w[j]) - dl) * p.scale);
          }
          if (MODE == ATT_DQ) {
            store_p16(P1, r, c0, ds);
          } else {
            store_p16(P1, r, c0, pe);
            store_p16(P2, r, c0, ds);
          }
        };
        tmem_ld_32x16(T_S + lane_addr + cbeg, va);
        tmem_ld_32x16(T_S2 + lane_addr + cbeg, wa);
#pragma unroll 1
        for (int c = 0; c < nch; c += 2) {
          tmem_wait_ld();
          if (c + 1 < nch) {
            tmem_ld_32x16(T_S + lane_addr + cbeg + (c + 1) * 16, vb);
            tmem_ld_32x16(T_S2 + lane_addr + cbeg + (c + 1) * 16, wb);
          }
          process(va, wa, cbeg + c * 16);
          if (c + 1 < nch) {
            tmem_wait_ld();
            if (c + 2 < nch) {
              tmem_ld_32x16(T_S + lane_addr + cbeg + (c + 2) * 16, va);
              tmem_ld_32x16(T_S2 + lane_addr + cbeg + (c + 2) * 16, wa);
            }
            process(vb, wb, cbeg + (c + 1) * 16);
          }
        }
      }
      if (sdbg) sdbg[3] = clock64();
      fence_async_smem();       // generic-proxy smem writes -> visible to the UMMA operand reads
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
      if (sdbg) sdbg[4] = clock64();

      mbar_wait(bar_o, ph, 21);
      tc_fence_after();
      if (sdbg) sdbg[5] = clock64();
      // ---- epilogue: FWD / DQ: this warp stores 32 of the 64 output columns; DKV: hf 0 -> dV, hf 1 -> dK
      {
        float inv = 1.f;
        if (MODE == ATT_FWD) {
          asm volatile("bar.sync 2, 256;" ::: "memory");   // both halves' partial sums are in smem
          const float tot = sum_s[r] + sum_s[128 + r];
          inv = 1.f / tot;
          if (hf == 0 && row_ok && p.lse) p.lse[((long long)b * H + h) * N + n] = max_s[r] * p.scale + logf(tot);
        }
        bf16* dst;
        uint32_t tsrc;
        int ncols;
        if (MODE == ATT_FWD) {
          dst = p.out + ((long long)b * N + n) * HD + (long long)h * ATT_DH + hf * 32;
          tsrc = T_OUT1 + hf * 32; ncols = 32;
        } else if (MODE == ATT_DQ) {
          dst = p.dqkv + (((long long)b * N + n) * 3 + 0) * HD + (long long)h * ATT_DH + hf * 32;
          tsrc = T_OUT1 + hf * 32; ncols = 32;
        } else {
          dst = p.dqkv + (((long long)b * N + n) * 3 + (hf == 0 ? 2 : 1)) * HD + (long long)h * ATT_DH;
          tsrc = hf == 0 ? T_OUT1 : T_OUT2; ncols = 64;
        }
        tsrc += lane_addr;
        for (int c0 = 0; c0 < ncols; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tsrc + c0, v);
          tmem_wait_ld();
          if (row_ok) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[8 * u + j]) * inv;
              V8<bf16>::store(dst + c0 + 8 * u, f);
            }
          }
        }
      }
      if (sdbg) sdbg[6] = clock64();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_free);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == CTRL) tmem_dealloc(tmem, 512);
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
static long long* g_attn_dbg = nullptr;
void attn_tc_set_debug(long long* buf) { g_attn_dbg = buf; }
long long* attn_tc_get_debug() { return g_attn_dbg; }

bool attn_tc_supported(int N, int dh, int dtype) {
  return dtype == NRV_BF16 && dh == ATT_DH && N >= 1 && N <= 208;
}

template <int MODE>
static int launch_att(const AttnParams& p, const CUtensorMap* maps, cudaStream_t st) {
  using L = AttSmem<MODE>;
  const int smem = L::total(p.NP);
  NRV_REQUIRE(smem <= 227 * 1024, "tcgen05 attention: %d bytes of shared memory needed (N=%d)", smem, p.N);
  NRV_CUDA(cudaFuncSetAttribute(attn_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = p.items < num_sms() ? p.items : num_sms();
  attn_tc_kernel<MODE><<<grid, ATT_THREADS, smem, st>>>(maps[0], maps[1], maps[2], maps[3], p);
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

static int make_maps(CUtensorMap* maps, const void* qkv, const void* dout, int B, int N, int H, int NP) {
  const uint64_t row_qkv = (uint64_t)3 * H * ATT_DH, row_o = (uint64_t)H * ATT_DH;
  int rc;
  rc = encode_tmap_3d(&maps[0], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, qkv, row_qkv, N, B, row_qkv * 2, row_qkv * 2 * N,
                      64, 128, 1, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = encode_tmap_3d(&maps[1], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, qkv, row_qkv, N, B, row_qkv * 2, row_qkv * 2 * N,
                      64, NP, 1, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  const void* d = dout ? dout : qkv;  // unused in FWD; keep the descriptors valid
  const uint64_t rd = dout ? row_o : row_qkv;
  rc = encode_tmap_3d(&maps[2], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, d, rd, N, B, rd * 2, rd * 2 * N, 64, 128, 1,
                      CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = encode_tmap_3d(&maps[3], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, d, rd, N, B, rd * 2, rd * 2 * N, 64, NP, 1,
                      CU_TENSOR_MAP_SWIZZLE_128B);
  return rc;
}

int attn_fwd_tc(const void* qkv, void* out, float* lse, int B, int N, int H, int dh, float scale, cudaStream_t st) {
  // second-generation forward (attention_fwd2.cu) unless the A/B switch or the timestamp hook asks for this one
  static const bool env_v1 = getenv("NRV_ATTN_V1") != nullptr;
  if (!env_v1) return attn_fwd_tc2(qkv, out, lse, B, N, H, dh, scale, st);
  NRV_REQUIRE(attn_tc_supported(N, dh, NRV_BF16), "tcgen05 attention: unsupported shape N=%d dh=%d", N, dh);
  NRV_REQUIRE(((uintptr_t)qkv % 16) == 0 && ((uintptr_t)out % 16) == 0, "tcgen05 attention: 16-byte alignment");
  AttnParams p{};
  p.B = B; p.N = N; p.H = H; p.NP = (N + 15) / 16 * 16;
  p.tiles = (N + 127) / 128; p.items = B * H;
  p.scale = scale; p.scale_log2e = scale * 1.4426950408889634f;
  p.out = (bf16*)out; p.lse = lse; p.dbg = g_attn_dbg;
  CUtensorMap maps[4];
  int rc = make_maps(maps, qkv, nullptr, B, N, H, p.NP);
  if (rc) return rc;
  return launch_att<ATT_FWD>(p, maps, st);
}

int attn_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* delta,
                int B, int N, int H, int dh, float scale, cudaStream_t st) {
  NRV_REQUIRE(attn_tc_supported(N, dh, NRV_BF16), "tcgen05 attention: unsupported shape N=%d dh=%d", N, dh);
  // second-generation fused backward (attention_bwd2.cu) unless the A/B switch asks for the two-pass kernels
  static const bool env_v1 = getenv("NRV_ATTN_V1") != nullptr;
  if (!env_v1) return attn_bwd_tc2(qkv, out, dout, lse, dqkv, B, N, H, dh, scale, st);
  NRV_REQUIRE(delta != nullptr, "tcgen05 attention backward needs a [B,H,N] fp32 scratch (delta)");
  AttnParams p{};
  p.B = B; p.N = N; p.H = H; p.NP = (N + 15) / 16 * 16;
  p.tiles = (N + 127) / 128; p.items = B * H;
  p.scale = scale; p.scale_log2e = scale * 1.4426950408889634f;
  p.o = (const bf16*)out; p.dout = (const bf16*)dout; p.dqkv = (bf16*)dqkv;
  p.lse = const_cast<float*>(lse); p.delta = delta; p.dbg = g_attn_dbg;
  CUtensorMap maps[4];
  int rc = make_maps(maps, qkv, dout, B, N, H, p.NP);
  if (rc) return rc;
  rc = launch_att<ATT_DQ>(p, maps, st);
  if (rc) return rc;
  return launch_att<ATT_DKV>(p, maps, st);
}

}  // namespace nrv
