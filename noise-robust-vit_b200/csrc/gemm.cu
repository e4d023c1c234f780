// Persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[M,N] = epilogue( alpha * sum_k A[m,k] * B[n,k] )
//
// * operands bf16 (or fp32 consumed as tf32 in check mode), fp32 accumulation in TMEM
// * A / B each either K-major (row-major [rows, K]) or MN-major (row-major [K, rows]) so the same
//   kernel serves forward (K,K), dX (K,MN) and dW (MN,MN) products of an nn.Linear without any
//   transposed copy in HBM (reference ops replaced: aten::mm / addmm behind every nn.Linear,
//   simple_vit.py:37-42,61-62,76 ; vit.py:41-47,105-111,237-242)
// * warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warps 2..5 = epilogue
// * TMEM double buffered (2 x BN fp32 columns) so tile i's epilogue overlaps tile i+1's mainloop
// * epilogue: TMEM -> regs -> swizzled smem transpose -> coalesced 16-byte global accesses with
//   fused bias / GELU / GELU' / residual / positional-embedding / token-row remap / fp32 split-K red
#include "common.cuh"
#include "nrvit_internal.h"

#include <mutex>
#include <stdlib.h>
#include <string.h>
#include <type_traits>
#include <vector>

namespace nrv {

constexpr int BM = 128;
constexpr int BK_BYTES = 128;  // one 128-byte swizzle row of K (64 bf16 / 32 tf32)
constexpr int NUM_EPI_WARPS = 8;
constexpr int EPI_WARP0 = 2;
constexpr int NUM_THREADS = 32 * (EPI_WARP0 + NUM_EPI_WARPS);
constexpr int STAGING_BYTES_PER_WARP = 4096;  // 32 rows x 128 bytes, 128B-swizzled (TMA store box)

struct GemmKernelParams {
  int M, N, K;
  int num_m_tiles, num_n_tiles, splits, kb_total, kb_per_split;
  // DUAL with an odd number of 256-row blocks: the last unit row is half dead (one MMA per K step).  Those `dual_half`
  // cheap units are dealt two at a time to the first `dual_rr` CTA pairs -- the ones the round-robin hands one unit more
  // than the rest -- so the longest pair does floor(units / pairs) full units instead of that plus a half (see unit_decode)
  int dual_half, dual_rr;
  int a_mn, b_mn;       // 1 = MN-major operand
  int tf32;             // 1 = fp32 operands through kind::tf32
  // descriptor increments (in 16-byte units) and strides (bytes)
  uint32_t a_kstep, b_kstep, a_lbo, b_lbo, a_sbo, b_sbo;
  uint32_t idesc;
  // epilogue
  int epi;              // NRV_EPI_*
  int tma_epi;          // 1: TMEM -> regs -> swizzled smem -> TMA store ; 0: generic path (row remap / atomics)
  int extra;            // tile-shaped epilogue input fetched by TMA: 0 none, 1 residual, 2 aux (GELU' pre-activation)
  float alpha;
  void* out; long long ldo;
  void* out2;           // GELU: pre-activation copy (ld = ldo)
  const float* bias;
  const void* residual; long long ldr;
  const void* aux; long long ldaux;  // DGELU: pre-activation (dtype of out)
  float* colsum;        // EPI_MUL (bf16): column sums of the stored tile, red.global.add
  const float* pos; int pos_rows_in, pos_rows_out, pos_row_off; long long ldpos;
  int out_f32;          // store fp32 instead of bf16 (check mode / logits)
  // LayerNorm folded into this product (A = the raw residual stream, B = row-centred gamma o W):  out = rstd_m acc_mn + c_n
  const double* ln_stats;  // [M][2] row sums (sum x, sum x^2) of A, or nullptr
  float ln_inv_dim, ln_eps;
  float* ln_mean_out;      // optional [M]: mean / rstd of every row, written by the blocks of column 0 (LayerNorm backward)
  float* ln_rstd_out;
  double* stats_out;       // optional [M][2]: += (sum, sum of squares) of the STORED output rows, fp64 reds (next LayerNorm)
  // PATCH instantiation (im2col fused into the operand loads, see PatchView in nrvit_internal.h).  The patch rows are
  // enumerated in "virtual" order: tile t of 128 rows = pm_npx neighbouring patches of one patch row (py) for pm_nb
  // consecutive images; t = (bg * pm_gh + py) * pm_nxg + xg; row r of the tile = (px = xg*pm_npx + r % pm_npx, b = bg*pm_nb + r / pm_npx).
  //   role 1 (forward):  M = virtual rows, A tiles come from the image through a 5-D map (p2, y, px, c, b)
  //   role 2 (dW):       K = virtual rows; a K block = half a tile; A = the gradient rows of the same tokens (3-D map
  //                      (d, token, b)), B = the image (5-D map with half the image count per box)
  int patch_role;
  int pm_npx, pm_nb, pm_nxg, pm_gh, pm_gw, pm_kbc, pm_rp1, pm_ph, pm_B, pm_tpi, pm_tok_off;
  // image tiles = pm_rp1 sub-tiles (one per patch row) of [rows x pm_line bytes] lines in the SWIZZLE_<pm_line>B layout:
  // descriptor layout type, 8-line group stride, sub-tile stride, and the offset of each of the 4 K steps of a K block
  uint32_t pm_layout, pm_sbo, pm_lbo, pm_koff[4];
};

// PAIR = two CTAs of a cluster run one cta_group::2 MMA of M = 256: each CTA stages its own 128 rows
// of A and HALF of the B tile, so a K block costs 32 KB of L2->SM traffic per CTA instead of 48 KB
// (the 128x256 single-CTA tile is L2-bandwidth bound on this part) and six stages fit.
//
// DUAL (pairs only) = one work unit is a 512 x 256 super tile: TWO M = 256 MMAs per K step, on two row blocks
// that share the B tile, each into its own 256-column accumulator (all 512 TMEM columns).  Measured on B200
// (profiles/r2b_*): the single-accumulator pair kernel pulls 41.6 B/clk/SM through the L2 -- the chip-wide L2
// delivery cap (~6300 B/clk) -- with the tensor pipe 70 % active; sharing B cuts the bytes per MAC by a quarter
// (48 KB per 2 x 512 MMA cycles instead of 32 KB per 512).  Price: no spare accumulator, so the next unit's MMAs
// wait until the epilogue warps have read the accumulators out (a bubble per unit; the dispatcher picks DUAL only
// where a unit has >= 24 K blocks).
template <int BN, bool PAIR, bool DUAL>
struct SmemLayout {
  static_assert(!DUAL || (PAIR && BN == 256), "DUAL needs the CTA-pair kernel with BN = 256");
  static constexpr int NSUB = DUAL ? 2 : 1;           // row blocks (accumulators) per work unit
  static constexpr int BN_CTA = PAIR ? BN / 2 : BN;   // B rows staged by one CTA
  static constexpr int A_BYTES = BM * BK_BYTES;       // per row block
  static constexpr int B_BYTES = BN_CTA * BK_BYTES;
  static constexpr int STAGE_BYTES = NSUB * A_BYTES + B_BYTES;
  static constexpr int STAGES = DUAL ? 4 : (PAIR ? 5 : ((BN == 256) ? 4 : 6));
  // epilogue staging: 128B-swizzled [32 rows x 128 B] tiles, the source of TMA stores and the landing
  // zone of TMA-loaded residual / pre-activation tiles.  The pair kernel double-buffers them per warp.
  static constexpr int NSTG = DUAL ? 1 : (PAIR ? 2 : 1);
  static constexpr int STAGING_OFF = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFF = STAGING_OFF + NUM_EPI_WARPS * NSTG * STAGING_BYTES_PER_WARP;
  // full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], extra[NUM_EPI_WARPS][2], tmem_ptr
  static constexpr int NBARS = 2 * STAGES + 4 + 2 * NUM_EPI_WARPS;
  static constexpr int TOTAL = BAR_OFF + NBARS * 8 + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for manual 1024-byte alignment
  static_assert(DYN_BYTES <= 232448, "shared memory budget");
};

// Work unit u -> (row tile, column tile, K split).  Natural order: column tiles fastest, then K splits, then rows.
// With half-dead units (see GemmKernelParams::dual_half) half unit h sits in slot k_h of pair c_h, everything else keeps
// its natural order among the full units.  `nclu` = CTA pairs of the launch.
__host__ __device__ __forceinline__ void unit_decode(int u, int nclu, int num_m_tiles, int num_n_tiles, int splits,
                                                     int dual_half, int dual_rr, int& m_t, int& n_t, int& s_t) {
  if (dual_half == 0) {
    n_t = u % num_n_tiles;
    s_t = (u / num_n_tiles) % splits;
    m_t = u / (num_n_tiles * splits);
    return;
  }
  int below = 0, hit = -1;
  for (int h = 0; h < dual_half; ++h) {
    const int c_h = (h >> 1) % dual_rr, k_h = (h & 1) + 2 * ((h >> 1) / dual_rr);
    const int u_h = c_h + k_h * nclu;
    if (u_h == u) hit = h;
    below += u_h < u ? 1 : 0;
  }
  s_t = 0;                      // (splits == 1 whenever dual_half != 0)
  if (hit >= 0) { m_t = num_m_tiles - 1; n_t = hit; return; }
  const int f = u - below;
  m_t = f / num_n_tiles;
  n_t = f % num_n_tiles;
}

// One [32 rows x NC columns] block of the accumulator, thread = row.  The fused epilogue math runs on
// registers; tile-shaped inputs (residual, GELU' pre-activation) arrive by TMA in the staging buffer and
// are combined IN PLACE (same thread, same 16-byte units), then the buffer leaves as one TMA store.
struct EpiBlock { int u, c; };   // work unit, column offset inside the warp's half tile

// LNX: the epilogue variants of the folded LayerNorm (apply the row statistics / emit them).  A separate instantiation of
// the whole kernel, because their extra live values push the common epilogues over the 168-register budget (measured: -5 %
// on every GEMM of the step when they were runtime branches of one kernel).
template <int NC, bool OUT_F32, bool LNX, bool PATCH, typename AfterLoad>
__device__ __forceinline__ void epi_math_and_store(const GemmKernelParams& p, const CUtensorMap* tmO,
                                                   const CUtensorMap* tmO2, uint32_t t_addr, uint8_t* stg_cur,
                                                   uint8_t* stg_alt, uint8_t* stg0, bool two_bufs, uint32_t extra_bar,
                                                   uint32_t extra_phase, int lane, int col0, int row0,
                                                   AfterLoad after_load) {
  using TO = typename std::conditional<OUT_F32, float, bf16>::type;
  constexpr int UNIT = 16 / (int)sizeof(TO);          // elements per 16-byte unit
  float x[NC];
  {
    // both 32-column loads in flight, one wait: under a running mainloop a TMEM round trip is slow (the MMAs'
    // accumulator traffic shares the port), so the epilogue pays for as few of them as possible.  The bias words
    // (one L1 hit per 4 columns, the same address in every lane) are requested while the TMEM loads are in flight:
    // issued after the wait, every FFMA2 below stalled on its own load (long_scoreboard, ncu r2b).
    uint32_t v[NC / 32][32];
#pragma unroll
    for (int h = 0; h < NC / 32; ++h) tmem_ld_32x32(t_addr + h * 32, v[h]);
    float4 bv[NC / 4];
#pragma unroll
    for (int j = 0; j < NC / 4; ++j) {
      bv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.bias != nullptr && col0 + 4 * j < p.N) bv[j] = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 4 * j));
    }
    float mu = 0.f, rs = 1.f;
    if (LNX && p.ln_stats != nullptr) {
      // row statistics of the LayerNorm folded into this GEMM (thread = row): requested with the bias words
      // fp64 sums: E[x^2] - mu^2 without cancellation, and cross-block accumulation order cannot show in the fp32 results
      double2 st = make_double2(0.0, 1.0);
      if (row0 + lane < p.M) st = __ldg(reinterpret_cast<const double2*>(p.ln_stats) + row0 + lane);
      const double mud = st.x * (double)p.ln_inv_dim;
      const double var = fmax(st.y * (double)p.ln_inv_dim - mud * mud, 0.0);
      mu = (float)mud;
      rs = rsqrtf((float)var + p.ln_eps);
      if (col0 == 0 && p.ln_mean_out != nullptr && row0 + lane < p.M) {
        p.ln_mean_out[row0 + lane] = mu;
        p.ln_rstd_out[row0 + lane] = rs;
      }
    }
    tmem_wait_ld();
    after_load();   // the accumulator values are in registers: the last block of a unit hands TMEM back here
    // x = alpha * acc + bias, two columns per issue slot
    // folded LayerNorm: B holds the row-centred gamma o W (rows sum to zero), so acc = sum_k (x_k - mu) gamma_k W_nk already
    // and LayerNorm(x) W^T + b = rstd * acc + c: the plain bias epilogue with a per-row alpha
    const uint64_t a2 = (LNX && p.ln_stats != nullptr) ? f2_pack(rs, rs) : f2_pack(p.alpha, p.alpha);
#pragma unroll
    for (int h = 0; h < NC / 32; ++h)
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = bv[h * 8 + j / 4];
        const int c = h * 32 + j;
        f2_unpack(f2_fma(f2_pack(__uint_as_float(v[h][j]), __uint_as_float(v[h][j + 1])), a2, f2_pack(b.x, b.y)), x[c], x[c + 1]);
        f2_unpack(f2_fma(f2_pack(__uint_as_float(v[h][j + 2]), __uint_as_float(v[h][j + 3])), a2, f2_pack(b.z, b.w)), x[c + 2], x[c + 3]);
      }
  }
  // patch embedding (forward): the warp's 32 rows are min(npx, 32) neighbouring patches of 32 / min(npx, 32) images; add the
  // positional row of this thread's token and let ONE 3-D TMA store (columns, tokens, images) scatter the block to its
  // token rows -- images beyond the batch are clipped by the tensor map
  int pt_tok0 = 0, pt_b0 = 0;
  if constexpr (PATCH) {
    const int tile = row0 / BM, r0 = row0 % BM;
    const int xg = tile % p.pm_nxg, py = (tile / p.pm_nxg) % p.pm_gh, bg = tile / (p.pm_nxg * p.pm_gh);
    const int w = p.pm_npx < 32 ? p.pm_npx : 32;                     // tokens per image inside the warp's 32 rows
    pt_tok0 = p.pm_tok_off + py * p.pm_gw + xg * p.pm_npx + (r0 % p.pm_npx);
    pt_b0 = bg * p.pm_nb + r0 / p.pm_npx;
    if (p.pos != nullptr) {
      const float* pp = p.pos + (long long)(pt_tok0 + lane % w) * p.ldpos + col0;
#pragma unroll
      for (int j = 0; j < NC; j += 4)
        if (col0 + j < p.N) {
          const float4 q0 = __ldg(reinterpret_cast<const float4*>(pp + j));
          x[j] += q0.x; x[j + 1] += q0.y; x[j + 2] += q0.z; x[j + 3] += q0.w;
        }
    }
  }
  // (sum, sum of squares) of this thread's row for the LayerNorm that consumes the output: one fp64 red pair per row and
  // block; columns beyond N are zero (TMA zero-fills B and the residual).  Taken from the fp32 values before the store
  // rounds them: the difference to the stored row is rounding noise (mean ~1e-4 sigma), and the consumer does not rely on
  // it -- the mean cancels through the zero-sum rows of its B operand, the statistics only set the scale.
  auto emit_stats = [&]() {
    if (!LNX || p.stats_out == nullptr) return;
    uint64_t s12 = f2_pack(0.f, 0.f), q12 = f2_pack(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < NC; j += 2) {
      const uint64_t v = f2_pack(x[j], x[j + 1]);
      s12 = f2_add(s12, v);
      q12 = f2_fma(v, v, q12);
    }
    float lo, hi, qlo, qhi;
    f2_unpack(s12, lo, hi);
    f2_unpack(q12, qlo, qhi);
    if (row0 + lane < p.M) {
      // fp64 reds: the within-block sums above are in a fixed order, the cross-block order is not, but its effect (1e-16
      // relative) disappears when the consumer rounds mean / rstd to fp32: the forward pass stays run-to-run reproducible
      double* o = p.stats_out + 2 * (long long)(row0 + lane);
      asm volatile("red.global.add.f64 [%0], %1;" ::"l"(o), "d"((double)(lo + hi)) : "memory");
      asm volatile("red.global.add.f64 [%0], %1;" ::"l"(o + 1), "d"((double)(qlo + qhi)) : "memory");
    }
  };
  auto write_tile = [&](uint8_t* stg) {
#pragma unroll
    for (int u = 0; u < NC / UNIT; ++u) {
      const uint32_t dst = smem_u32(stg) + lane * 128 + ((u ^ (lane & 7)) << 4);
      if (OUT_F32) {
        sts128f(dst, x[4 * u], x[4 * u + 1], x[4 * u + 2], x[4 * u + 3]);
      } else {
        sts128(dst, pack_bf16(x[8 * u], x[8 * u + 1]), pack_bf16(x[8 * u + 2], x[8 * u + 3]),
               pack_bf16(x[8 * u + 4], x[8 * u + 5]), pack_bf16(x[8 * u + 6], x[8 * u + 7]));
      }
    }
  };
  auto send_tile = [&](const CUtensorMap* tm, uint8_t* stg) {
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      if constexpr (PATCH) tma_store_3d(tm, smem_u32(stg), col0, pt_tok0, pt_b0);
      else tma_store_2d(tm, smem_u32(stg), col0, row0);
      tma_store_commit();
    }
  };
  if (p.extra != 0) {
    // ---- residual / pre-activation tile is (being) loaded into stg_cur by TMA
    mbar_wait(extra_bar, extra_phase, 5);
    if (!OUT_F32 && p.extra == 3) {
      // EPI_MUL in bf16: round the accumulator to bf16 and multiply the packed pairs by the stored factor (one
      // HMUL2.BF16 per two elements, no unpacking) -- the two roundings a bf16 autocast graph performs here
#pragma unroll
      for (int u = 0; u < NC / 8; ++u) {
        const uint32_t src = smem_u32(stg_cur) + lane * 128 + ((u ^ (lane & 7)) << 4);
        const uint4 q = lds128(src);
        sts128(src, bf2_mul(pack_bf16(x[8 * u], x[8 * u + 1]), q.x), bf2_mul(pack_bf16(x[8 * u + 2], x[8 * u + 3]), q.y),
               bf2_mul(pack_bf16(x[8 * u + 4], x[8 * u + 5]), q.z), bf2_mul(pack_bf16(x[8 * u + 6], x[8 * u + 7]), q.w));
      }
      send_tile(tmO, stg_cur);
      if (p.colsum != nullptr) {
        // Column sums of the tile as stored (rows beyond M are zero: TMA zero-fills A and the factor tile).  Lane l owns
        // the bf16 pair of columns 2l, 2l+1: word (l & 3) of 16-byte unit (l >> 2) ^ (row & 7) of every 128-byte row --
        // the 32 lanes read the 32 words of one row, conflict-free.  The TMA store reads the same tile concurrently.
        uint64_t acc2 = f2_pack(0.f, 0.f);
        const uint32_t base = smem_u32(stg_cur) + (lane & 3) * 4;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
          uint32_t w;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(base + r * 128 + ((((lane >> 2) ^ (r & 7))) << 4)) : "memory");
          acc2 = f2_add(acc2, f2_pack(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)));
        }
        float s0, s1;
        f2_unpack(acc2, s0, s1);
        const int c = col0 + 2 * lane;
        if (c < p.N)
          asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p.colsum + c), "f"(s0), "f"(s1) : "memory");
      }
      return;
    }
#pragma unroll
    for (int u = 0; u < NC / UNIT; ++u) {
      const uint32_t src = smem_u32(stg_cur) + lane * 128 + ((u ^ (lane & 7)) << 4);
      float e[UNIT];
      if (OUT_F32) {
        const float4 f = lds128f(src);
        e[0] = f.x; e[1] = f.y; e[2] = f.z; e[3] = f.w;
      } else {
        const uint4 q = lds128(src);
        const float2 a = unpack_bf16(q.x), b = unpack_bf16(q.y), c = unpack_bf16(q.z), d = unpack_bf16(q.w);
        e[0] = a.x; e[1] = a.y; e[2] = b.x; e[3] = b.y;
        if (UNIT == 8) { e[4 % UNIT] = c.x; e[5 % UNIT] = c.y; e[6 % UNIT] = d.x; e[7 % UNIT] = d.y; }
      }
#pragma unroll
      for (int j = 0; j < UNIT; ++j) {
        if (p.extra == 2) x[u * UNIT + j] *= dgelu_erf(e[j]);
        else if (p.extra == 3) x[u * UNIT + j] *= e[j];
        else x[u * UNIT + j] += e[j];
      }
    }
    emit_stats();
    write_tile(stg_cur);         // in place: every thread rewrites exactly the units it read
    send_tile(tmO, stg_cur);
    return;
  }
  if (p.epi == NRV_EPI_GELU_GRAD) {
    // out = gelu(u) and out2 = gelu'(u) from the same cdf / density: the backward GEMM then only multiplies
    // (EPI_MUL) instead of re-evaluating erf and exp on its own epilogue.  gelu' leaves through buffer 0,
    // packed unit by unit so only one block of fp32 values is live.
    uint8_t* bg = two_bufs ? stg0 : stg_cur;
    uint8_t* bh = two_bufs ? stg0 + STAGING_BYTES_PER_WARP : stg_cur;
    if (lane == 0) { if (two_bufs) tma_store_wait_read<1>(); else tma_store_wait_read<0>(); }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < NC / UNIT; ++u) {
      float g[UNIT];
      if (OUT_F32) {                 // check mode: Abramowitz-Stegun erf (1.5e-7)
#pragma unroll
        for (int j = 0; j < UNIT; ++j) {
          const float v = x[u * UNIT + j];
          float cdf, e;
          gelu_parts(v, cdf, e);
          x[u * UNIT + j] = v * cdf;
          g[j] = fmaf(v * 0.39894228040143267794f, e, cdf);
        }
      } else {                       // production: packed sigmoid form, two elements per issue slot
#pragma unroll
        for (int j = 0; j < UNIT; j += 2)
          gelu_sig_pair(x[u * UNIT + j], x[u * UNIT + j + 1], true, x[u * UNIT + j], x[u * UNIT + j + 1], g[j], g[j + 1]);
      }
      const uint32_t dst = smem_u32(bg) + lane * 128 + ((u ^ (lane & 7)) << 4);
      if (OUT_F32) {
        sts128f(dst, g[0], g[1], g[2], g[3]);
      } else {
        sts128(dst, pack_bf16(g[0], g[1]), pack_bf16(g[2], g[3]), pack_bf16(g[4 % UNIT], g[5 % UNIT]), pack_bf16(g[6 % UNIT], g[7 % UNIT]));
      }
    }
    send_tile(tmO2, bg);
    if (lane == 0) { if (two_bufs) tma_store_wait_read<1>(); else tma_store_wait_read<0>(); }
    __syncwarp();
    write_tile(bh);
    send_tile(tmO, bh);
    return;
  }
  if (p.epi == NRV_EPI_GELU) {
    // two outputs per block.  With two staging buffers u always goes through buffer 0 and gelu(u) through
    // buffer 1, so each write only has to wait for the store issued two groups earlier.
    uint8_t* bu = two_bufs ? stg0 : stg_cur;
    uint8_t* bh = two_bufs ? stg0 + STAGING_BYTES_PER_WARP : stg_cur;
    if (p.out2 != nullptr) {     // pre-activation u, kept for backward
      if (lane == 0) { if (two_bufs) tma_store_wait_read<1>(); else tma_store_wait_read<0>(); }
      __syncwarp();
      write_tile(bu);
      send_tile(tmO2, bu);
    }
    if (OUT_F32) {
#pragma unroll
      for (int j = 0; j < NC; ++j) x[j] = gelu_erf(x[j]);
    } else {
#pragma unroll
      for (int j = 0; j < NC; j += 2) {
        float g0, g1;
        gelu_sig_pair(x[j], x[j + 1], false, x[j], x[j + 1], g0, g1);
      }
    }
    if (lane == 0) { if (two_bufs && p.out2 != nullptr) tma_store_wait_read<1>(); else tma_store_wait_read<0>(); }
    __syncwarp();
    write_tile(bh);
    send_tile(tmO, bh);
    return;
  }
  // plain store: the previous store that used this buffer must have finished reading it
  emit_stats();
  if (lane == 0) { if (two_bufs) tma_store_wait_read<1>(); else tma_store_wait_read<0>(); }
  __syncwarp();
  write_tile(stg_cur);
  send_tile(tmO, stg_cur);
}

template <int BN, bool PAIR, bool DUAL, bool LNX, bool PATCH = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmO2,
            const __grid_constant__ CUtensorMap tmX, const GemmKernelParams p) {
  using L = SmemLayout<BN, PAIR, DUAL>;
  constexpr int NSUB = L::NSUB;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;               // 0 = leader (issues the MMAs)
  const int cid = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // work-loop index of this CTA / pair
  const int nclu = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int BM_SUB = PAIR ? 2 * BM : BM;                         // rows of one accumulator (row block)
  constexpr int BM_UNIT = NSUB * BM_SUB;                             // rows of one work unit
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_base = sbase + L::BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (L::STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * L::STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * L::STAGES + 2 + s); };
  auto extra_bar = [&](int ew, int s) { return bar_base + 8u * (2 * L::STAGES + 4 + 2 * ew + s); };
  volatile uint32_t* tmem_ptr_smem =
      reinterpret_cast<volatile uint32_t*>(smem + L::BAR_OFF + L::NBARS * 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_epi) { tma_prefetch_desc(&tmO); tma_prefetch_desc(&tmO2); tma_prefetch_desc(&tmX); }
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < L::STAGES; ++s) {
        mbar_init(full_bar(s), 1);
        mbar_init(empty_bar(s), 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(tfull_bar(s), 1);
        // DUAL: accumulator s belongs to epilogue warps 4s .. 4s+3 of both CTAs; otherwise all 8 warps read both halves
        mbar_init(tempty_bar(s), (PAIR ? 2 : 1) * (DUAL ? NUM_EPI_WARPS / 2 : NUM_EPI_WARPS));
      }
      for (int w = 0; w < NUM_EPI_WARPS; ++w) {
        mbar_init(extra_bar(w, 0), 1);
        mbar_init(extra_bar(w, 1), 1);
      }
      fence_barrier_init();
    }
    __syncwarp();
    if (PAIR) tmem_alloc_pair(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), 2 * BN);
    else tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), 2 * BN);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();   // peer barriers initialised before any remote arrive / TMA credit
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // PDL: everything above ran under the tail of the previous kernel in the stream; its results are needed from here on.
  // The next kernel may begin its own prologue as soon as SMs free up.
  pdl_wait();
  pdl_launch_dependents();

  const int total_units = p.num_m_tiles * p.num_n_tiles * p.splits;
  auto decode = [&](int u, int& m_t, int& n_t, int& s_t) {
    unit_decode(u, nclu, p.num_m_tiles, p.num_n_tiles, p.splits, DUAL ? p.dual_half : 0, p.dual_rr, m_t, n_t, s_t);
  };
  // DUAL: the second row block of the last unit row may lie entirely beyond M (odd number of 256-row blocks):
  // nothing is loaded, multiplied or stored for it (a pair-uniform decision)
  auto sub_live = [&](int m_t, int sub) { return !DUAL || sub == 0 || m_t * BM_UNIT + sub * BM_SUB < p.M; };

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const int kelems = p.tf32 ? 32 : 64;  // K elements per 128-byte row
      auto load = [&](uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
        if (PAIR) tma_load_2d_pair(dst, tm, bar, c0, c1);   // bytes credited to the leader's barrier
        else tma_load_2d(dst, tm, bar, c0, c1);
      };
      for (int u = cid; u < total_units; u += nclu) {
        int m_t, n_t, s_t;
        decode(u, m_t, n_t, s_t);
        const int n0 = n_t * BN + (int)rank * L::BN_CTA;
        const int kb0 = s_t * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        const int nlive = (NSUB == 2 && sub_live(m_t, 1)) ? 2 : 1;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1, 1);
          const uint32_t sa = sbase + stage * L::STAGE_BYTES;
          const uint32_t sb = sa + NSUB * L::A_BYTES;
          if (!PAIR || rank == 0)
            mbar_arrive_expect_tx(full_bar(stage), (PAIR ? 2 : 1) * (nlive * L::A_BYTES + L::B_BYTES));
          const int k0 = kb * kelems;
          if constexpr (PATCH) {
            // im2col through TMA: the operand tile is gathered from the image (and, for dW, from the gradient rows of
            // the same tokens) by multi-dimensional boxes; the smem tiles have the ordinary layouts
            const int tile = p.patch_role == 1 ? (m_t * BM_UNIT + (int)rank * BM) / BM : (kb >> 1);
            const int xg = tile % p.pm_nxg, py = (tile / p.pm_nxg) % p.pm_gh, bg = tile / (p.pm_nxg * p.pm_gh);
            if (p.patch_role == 1) {
              const int ch = kb / p.pm_kbc, y = py * p.pm_ph + (kb % p.pm_kbc) * p.pm_rp1;
              tma_load_5d_pair(sa, &tmA, full_bar(stage), 0, xg * p.pm_npx, bg * p.pm_nb, y, ch);
              load(sb, &tmB, full_bar(stage), k0, n0);
            } else {
              const int b0 = bg * p.pm_nb + (kb & 1) * (p.pm_nb >> 1);
              const int tok0 = p.pm_tok_off + py * p.pm_gw + xg * p.pm_npx;
              const int m0 = m_t * BM_UNIT + (int)rank * BM;
#pragma unroll 1
              for (int c = 0; c < BM * 2 / 128; ++c)
                tma_load_3d_pair(sa + c * (BK_BYTES * 64), &tmA, full_bar(stage), m0 + c * 64, tok0, b0);
#pragma unroll 1
              for (int c = 0; c < L::BN_CTA * 2 / 128; ++c) {
                const int kbn = (n0 + c * 64) >> 6;           // 64 patch elements = one (channel, patch-row group)
                const int ch = kbn / p.pm_kbc, y = py * p.pm_ph + (kbn % p.pm_kbc) * p.pm_rp1;
                tma_load_5d_pair(sb + c * (BK_BYTES * 64), &tmB, full_bar(stage), 0, xg * p.pm_npx, b0, y, ch);
              }
            }
            if (++stage == L::STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
#pragma unroll
          for (int sub = 0; sub < NSUB; ++sub) {
            if (sub >= nlive) break;
            const int m0 = m_t * BM_UNIT + sub * BM_SUB + (int)rank * BM;
            const uint32_t dst = sa + sub * L::A_BYTES;
            if (!p.a_mn) {
              load(dst, &tmA, full_bar(stage), k0, m0);
            } else {
              // box {kelems of M, kelems.. rows of K}: one 128-byte-wide M chunk per issue
#pragma unroll 1
              for (int c = 0; c < BM * (p.tf32 ? 4 : 2) / 128; ++c)
                load(dst + c * (BK_BYTES * kelems), &tmA, full_bar(stage), m0 + c * kelems, k0);
            }
          }
          if (!p.b_mn) {
            load(sb, &tmB, full_bar(stage), k0, n0);
          } else {
#pragma unroll 1
            for (int c = 0; c < L::BN_CTA * (p.tf32 ? 4 : 2) / 128; ++c)
              load(sb + c * (BK_BYTES * kelems), &tmB, full_bar(stage), n0 + c * kelems, k0);
          }
          if (++stage == L::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA only when paired) ==========
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    if (!PAIR || rank == 0)
    for (int u = cid; u < total_units; u += nclu, ++it) {
      int m_t, n_t, s_t;
      decode(u, m_t, n_t, s_t);
      (void)n_t;
      const int kb0 = s_t * p.kb_per_split;
      const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
      // single accumulator per unit: stages alternate, tile i+1 accumulates while tile i is read out.
      // DUAL: both accumulators belong to this unit and must have been read out by the previous unit's epilogue.
      const int as = DUAL ? 0 : (it & 1);
      const uint32_t aphase = DUAL ? (uint32_t)(it & 1) : (uint32_t)((it >> 1) & 1);
      const int nlive = (NSUB == 2 && sub_live(m_t, 1)) ? 2 : 1;
      mbar_wait(tempty_bar(as), aphase ^ 1, 2);
      if (DUAL) mbar_wait(tempty_bar(1), aphase ^ 1, 2);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar(stage), phase, 3);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = sbase + stage * L::STAGE_BYTES;
          const uint32_t sb = sa + NSUB * L::A_BYTES;
          uint64_t bdesc = make_smem_desc_sw128(sb, p.b_lbo, p.b_sbo);
          if (PATCH && p.patch_role == 2) bdesc = make_smem_desc(sb, p.pm_lbo, p.pm_sbo, p.pm_layout);
#pragma unroll
          for (int sub = 0; sub < NSUB; ++sub) {
            if (sub >= nlive) break;
            uint64_t adesc = make_smem_desc_sw128(sa + sub * L::A_BYTES, p.a_lbo, p.a_sbo);
            if (PATCH && p.patch_role == 1) adesc = make_smem_desc(sa, 16, p.pm_sbo, p.pm_layout);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t acc = (kb > kb0 || k > 0) ? 1u : 0u;
              uint64_t ad = adesc + (uint64_t)(k * p.a_kstep), bd = bdesc + (uint64_t)(k * p.b_kstep);
              if (PATCH) {   // the image operand's K steps walk the sub-tiles
                if (p.patch_role == 1) ad = adesc + (uint64_t)(p.pm_koff[k] >> 4);
                else bd = bdesc + (uint64_t)(p.pm_koff[k] >> 4);
              }
              if (PAIR) {
                if (p.tf32) umma_tf32_pair(d_tmem + sub * BN, ad, bd, p.idesc, acc);
                else umma_bf16_pair(d_tmem + sub * BN, ad, bd, p.idesc, acc);
              } else {
                if (p.tf32) umma_tf32(d_tmem, ad, bd, p.idesc, acc);
                else umma_bf16(d_tmem, ad, bd, p.idesc, acc);
              }
            }
          }
          // smem slot free (in both CTAs when paired) once these MMAs retire
          if (PAIR) umma_commit_pair(empty_bar(stage), 3); else umma_commit(empty_bar(stage));
          if (kb == kb1 - 1) {                          // accumulator(s) complete
            if (PAIR) umma_commit_pair(tfull_bar(as), 3); else umma_commit(tfull_bar(as));
            if (DUAL) umma_commit_pair(tfull_bar(1), 3);
          }
        }
        __syncwarp();
        if (++stage == L::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================================== epilogue =========================================
    // 8 warps: TMEM lane quarter q = warp % 4 (hardware rule).  Single accumulator: column half = (warp - 2) / 4 of
    // the unit's accumulator.  DUAL: warps 2..5 own row block 0 (all BN columns), warps 6..9 row block 1.
    const int ew = warp - EPI_WARP0;
    const int q = warp & 3;
    const int grp = ew >> 2;                       // column half, or row block when DUAL
    constexpr int WARP_N = DUAL ? BN : BN / 2;     // accumulator columns one warp reads out per unit
    const int col_off = DUAL ? 0 : grp * WARP_N;   // first column (within the tile) of this warp
    const int row_off = (DUAL ? grp * BM_SUB : 0) + (int)rank * BM + q * 32;   // first row (within the unit)
    uint8_t* stg = smem + L::STAGING_OFF + ew * L::NSTG * STAGING_BYTES_PER_WARP;
    // ---- block enumeration for the TMA path: blocks of NCB columns of this warp's columns, dead
    //      blocks (beyond N, or a dead row block) skipped; `gb` counts live blocks of this warp over the whole kernel
    const int NCB = p.out_f32 ? 32 : 64;
    auto block_coords = [&](int u, int c, int& row0, int& col0) {
      int m_t, n_t, s_t;
      decode(u, m_t, n_t, s_t);
      row0 = m_t * BM_UNIT + row_off;
      col0 = n_t * BN + col_off + c;
    };
    auto unit_live = [&](int u) {
      int m_t, n_t, s_t;
      decode(u, m_t, n_t, s_t);
      return sub_live(m_t, DUAL ? grp : 0);
    };
    auto next_live = [&](int& u, int& c) -> bool {   // advance (u, c) to the next live block
      for (;;) {
        c += NCB;
        if (c >= WARP_N) { c = 0; u += nclu; }
        if (u >= total_units) return false;
        int r0, c0;
        block_coords(u, c, r0, c0);
        if (c0 < p.N && unit_live(u)) return true;
      }
    };
    auto issue_extra = [&](int u, int c, int gb) {   // TMA-load the residual / pre-activation tile of block gb
      if (lane == 0) {
        const int sidx = L::NSTG == 2 ? (gb & 1) : 0;
        // the buffer's previous TMA store must have finished READING it
        if (L::NSTG == 2) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
        int r0, c0;
        block_coords(u, c, r0, c0);
        const uint32_t bar = extra_bar(ew, sidx);
        mbar_arrive_expect_tx(bar, STAGING_BYTES_PER_WARP);
        tma_load_2d(smem_u32(stg + sidx * STAGING_BYTES_PER_WARP), &tmX, bar, c0, r0);
      }
      __syncwarp();
    };
    int gb = 0;                    // live blocks processed so far by this warp
    int pu = cid, pc = -NCB;       // prefetch cursor
    bool pre_ok = false;
    if (p.tma_epi && p.extra != 0) {
      pre_ok = next_live(pu, pc);
      if (pre_ok) issue_extra(pu, pc, 0);
    }
    int it = 0;
    for (int u = cid; u < total_units; u += nclu, ++it) {
      int m_t, n_t, s_t;
      decode(u, m_t, n_t, s_t);
      const int n0 = n_t * BN;
      const int as = DUAL ? grp : (it & 1);
      const uint32_t aphase = DUAL ? (uint32_t)(it & 1) : (uint32_t)((it >> 1) & 1);
      mbar_wait(tfull_bar(as), aphase, 4);
      tc_fence_after();
      const uint32_t t_row = tmem_base + as * BN + col_off + ((uint32_t)(q * 32) << 16);
      const int row0 = m_t * BM_UNIT + row_off;
      const int cbase = n0 + col_off;
      auto release_tmem = [&]() {
        // all TMEM reads of this accumulator (by this warp) are done: hand it back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_cluster(mapa_shared(tempty_bar(as), 0));   // the leader's barrier
          else mbar_arrive(tempty_bar(as));
        }
      };
      if (DUAL && !unit_live(u)) { release_tmem(); continue; }   // dead row block: nothing was accumulated
      if (p.tma_epi) {
#pragma unroll 1
        for (int c = 0; c < WARP_N; c += NCB) {
          if (cbase + c < p.N) {
            const int sidx = L::NSTG == 2 ? (gb & 1) : 0;
            uint8_t* cur = stg + sidx * STAGING_BYTES_PER_WARP;
            uint8_t* alt = stg + (L::NSTG == 2 ? (sidx ^ 1) : 0) * STAGING_BYTES_PER_WARP;
            const uint32_t xbar = extra_bar(ew, sidx);
            const uint32_t xph = L::NSTG == 2 ? ((gb >> 1) & 1) : (gb & 1);
            // last live block of this warp's columns: release the accumulator as soon as it has been read, so the
            // MMA warp can start the unit after next while this block's math and stores are still running
            const bool last_live = (c + NCB >= WARP_N) || (cbase + c + NCB >= p.N);
            auto after_load = [&]() { if (last_live) release_tmem(); };
            if (p.out_f32)
              epi_math_and_store<32, true, LNX, false>(p, &tmO, &tmO2, t_row + c, cur, alt, stg, L::NSTG == 2, xbar, xph, lane, cbase + c, row0, after_load);
            else
              epi_math_and_store<64, false, LNX, PATCH>(p, &tmO, &tmO2, t_row + c, cur, alt, stg, L::NSTG == 2, xbar, xph, lane, cbase + c, row0, after_load);
            ++gb;
            if (p.extra != 0) {
              // fetch the NEXT block's residual / pre-activation tile: its buffer was last read by the store
              // of block gb-2 (two buffers) or of this block (one buffer); issue_extra waits for exactly that
              pre_ok = next_live(pu, pc);
              if (pre_ok) issue_extra(pu, pc, gb);
            }
          }
          else if (c == 0) release_tmem();   // no live block in this warp's columns (N edge): nothing to read
        }
      } else {
        // ---- generic path: fp32 transpose through smem, 4 columns per thread (row remap, pos table,
        //      split-K atomics).  32-column chunks.
        float* stf = reinterpret_cast<float*>(stg);
        const int rsub = lane >> 3;  // row within a 4-row group
        const int cj = lane & 7;     // 4-column unit within the 32-column chunk
#pragma unroll 1
        for (int c = 0; c < WARP_N; c += 32) {
          const int col0 = cbase + c;
          const bool live = col0 < p.N;  // warp-uniform
          if (live) {
            uint32_t v[32];
            tmem_ld_32x32(t_row + c, v);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts128(smem_u32(stf + lane * 32 + ((j ^ (lane & 7)) << 2)), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          if (c + 32 >= WARP_N) release_tmem();
          if (!live) continue;
          __syncwarp();
          const int col = col0 + cj * 4;
          const bool col_ok = col < p.N;  // N % 8 == 0
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.bias != nullptr && col_ok) b4 = *reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll 2
          for (int r4 = 0; r4 < 8; ++r4) {
            const int r = r4 * 4 + rsub;
            const long long grow = (long long)row0 + r;
            if (grow >= p.M || !col_ok) continue;
            float4 f = lds128f(smem_u32(stf + r * 32 + ((cj ^ (r & 7)) << 2)));
            f.x = fmaf(f.x, p.alpha, b4.x); f.y = fmaf(f.y, p.alpha, b4.y);
            f.z = fmaf(f.z, p.alpha, b4.z); f.w = fmaf(f.w, p.alpha, b4.w);
            long long orow = grow;
            if (PATCH && p.patch_role == 1) {
              // virtual patch row -> token row (b, tok_off + py*gw + px); rows of images beyond the batch are dropped
              const int tile = (int)(grow / BM), rin = (int)(grow % BM);
              const int xg = tile % p.pm_nxg, py = (tile / p.pm_nxg) % p.pm_gh, bg = tile / (p.pm_nxg * p.pm_gh);
              const int b = bg * p.pm_nb + rin / p.pm_npx;
              if (b >= p.pm_B) continue;
              const int pr = p.pm_tok_off + py * p.pm_gw + xg * p.pm_npx + rin % p.pm_npx;
              orow = (long long)b * p.pm_tpi + pr;
              if (p.pos != nullptr) {
                const float4 q0 = *reinterpret_cast<const float4*>(p.pos + (long long)pr * p.ldpos + col);
                f.x += q0.x; f.y += q0.y; f.z += q0.z; f.w += q0.w;
              }
            } else if (p.pos_rows_in > 0) {
              // patch-embed: GEMM row (b, patch) -> token row (b, patch + off); add pos-emb row
              const long long b = grow / p.pos_rows_in;
              const int pr = (int)(grow - b * p.pos_rows_in) + p.pos_row_off;
              orow = b * p.pos_rows_out + pr;
              if (p.pos != nullptr) {
                const float4 q0 = *reinterpret_cast<const float4*>(p.pos + (long long)pr * p.ldpos + col);
                f.x += q0.x; f.y += q0.y; f.z += q0.z; f.w += q0.w;
              }
            }
            if (p.epi == NRV_EPI_ATOMIC_F32) {
              float* o = reinterpret_cast<float*>(p.out) + orow * p.ldo + col;
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(f.x), "f"(f.y),
                           "f"(f.z), "f"(f.w) : "memory");
            } else if (p.out_f32) {
              *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + orow * p.ldo + col) = f;
            } else {
              *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.out) + orow * p.ldo + col) =
                  make_uint2(pack_bf16(f.x, f.y), pack_bf16(f.z, f.w));
            }
          }
          __syncwarp();  // staging is overwritten by the next chunk
        }
      }
    }
    if (p.tma_epi && lane == 0) tma_store_wait<0>();  // smem must outlive the last bulk store
  }

  // ---- teardown ----
  tc_fence_before();
  if (PAIR) cluster_sync_all();   // the peer may still read this CTA's smem / arrive on its barriers
  else __syncthreads();
  if (warp == 1) {
    if (PAIR) tmem_dealloc_pair(tmem_base, 2 * BN);
    else tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ----------------------------------------------------------------------------------------------
// check mode (NRV_F32): 3xTF32 operand split.  kind::tf32 reads only the top 19 bits of an fp32
// operand, so x = hi + lo with hi = x & ~0x1fff (exact in TF32) and lo = x - hi gives
//   A*B ~= A_hi*B_hi + A_hi*B_lo + A_lo*B_hi      (error ~2^-21 relative)
// which is ONE K-major GEMM over K' = 3*Kp on   A' = [A_hi | A_hi | A_lo],  B' = [B_hi | B_lo | B_hi].
// The pre-pass also absorbs MN-major operands (it is a copy anyway), so the tcgen05 mainloop only
// ever sees K-major TF32 tiles.  dst: [rows, 3*Kp] fp32, Kp = K rounded up to 4, zero padded.
// ----------------------------------------------------------------------------------------------
__global__ void split3_kernel(const float* __restrict__ src, long long ld, int mn_major, int rows,
                              int K, int Kp, int is_b, float* __restrict__ dst) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  // load a 32(rows) x 32(k) tile into tile[r][k], coalesced along the contiguous source dimension
  for (int i = ty; i < 32; i += 8) {
    if (!mn_major) {
      const int r = r0 + i, k = k0 + tx;
      tile[i][tx] = (r < rows && k < K) ? src[(long long)r * ld + k] : 0.f;
    } else {
      const int k = k0 + i, r = r0 + tx;
      tile[tx][i] = (r < rows && k < K) ? src[(long long)k * ld + r] : 0.f;
    }
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, k = k0 + tx;
    if (r < rows && k < Kp) {
      const float x = tile[i][tx];
      const float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
      const float lo = x - hi;
      float* d = dst + (long long)r * (3 * Kp) + k;
      d[0] = hi;
      d[Kp] = is_b ? lo : hi;
      d[2 * Kp] = is_b ? hi : lo;
    }
  }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
// ---- optional per-launch timing (bench.py's roofline: CUDA events on the launching stream) --------
struct GemmTiming {
  std::mutex mu;
  bool enabled = false;
  std::vector<cudaEvent_t> ev;  // pairs
  std::vector<double> flops;
  std::vector<long long> shape;  // M, N, K, epi per launch
  size_t used = 0;
};
static GemmTiming g_timing;

static bool timing_begin(cudaEvent_t* e0, cudaEvent_t* e1, double flops, const nrv_gemm_desc* d = nullptr) {
  if (!g_timing.enabled) return false;
  std::lock_guard<std::mutex> lk(g_timing.mu);
  if (g_timing.used * 2 + 2 > g_timing.ev.size()) {
    cudaEvent_t a, b;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return false;
    g_timing.ev.push_back(a);
    g_timing.ev.push_back(b);
  }
  *e0 = g_timing.ev[g_timing.used * 2];
  *e1 = g_timing.ev[g_timing.used * 2 + 1];
  if (g_timing.flops.size() <= g_timing.used) g_timing.flops.push_back(flops);
  else g_timing.flops[g_timing.used] = flops;
  if (g_timing.shape.size() < (g_timing.used + 1) * 4) g_timing.shape.resize((g_timing.used + 1) * 4);
  if (d) {
    long long* sh = &g_timing.shape[g_timing.used * 4];
    sh[0] = d->M; sh[1] = d->N; sh[2] = d->K; sh[3] = d->epi | (d->a_layout << 4) | (d->b_layout << 5);
  }
  ++g_timing.used;
  return true;
}

void gemm_timing_enable(int on) {
  std::lock_guard<std::mutex> lk(g_timing.mu);
  g_timing.enabled = on != 0;
  if (on) g_timing.used = 0;
}

// per-launch records: out[i] = {M, N, K, epi|layouts, microseconds}; returns the number written
int gemm_timing_detail(long long* out, int max_records) {
  std::lock_guard<std::mutex> lk(g_timing.mu);
  int n = 0;
  for (size_t i = 0; i < g_timing.used && n < max_records; ++i, ++n) {
    float dt = 0.f;
    if (cudaEventSynchronize(g_timing.ev[2 * i + 1]) != cudaSuccess) break;
    cudaEventElapsedTime(&dt, g_timing.ev[2 * i], g_timing.ev[2 * i + 1]);
    for (int j = 0; j < 4; ++j) out[n * 5 + j] = g_timing.shape[i * 4 + j];
    out[n * 5 + 4] = (long long)(dt * 1000.0f);
  }
  return n;
}

int gemm_timing_read(double* ms, double* flops, long long* launches) {
  std::lock_guard<std::mutex> lk(g_timing.mu);
  double t = 0.0, f = 0.0;
  for (size_t i = 0; i < g_timing.used; ++i) {
    float dt = 0.f;
    NRV_CUDA(cudaEventSynchronize(g_timing.ev[2 * i + 1]));
    NRV_CUDA(cudaEventElapsedTime(&dt, g_timing.ev[2 * i], g_timing.ev[2 * i + 1]));
    t += dt;
    f += g_timing.flops[i];
  }
  if (ms) *ms = t;
  if (flops) *flops = f;
  if (launches) *launches = (long long)g_timing.used;
  return NRV_OK;
}

template <int BN, bool PAIR, bool DUAL, bool LNX, bool PATCH = false>
static int launch(const nrv_gemm_desc* d, const GemmKernelParams& kp, const CUtensorMap& ta,
                  const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& to2, const CUtensorMap& tx,
                  int grid, cudaStream_t stream) {
  using L = SmemLayout<BN, PAIR, DUAL>;
  static bool attr_set = false;
  if (!attr_set) {
    NRV_CUDA(cudaFuncSetAttribute(gemm_kernel<BN, PAIR, DUAL, LNX, PATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  L::DYN_BYTES));
    attr_set = true;
  }
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  const bool timed = timing_begin(&ev0, &ev1, 2.0 * (double)d->M * (double)d->N * (double)d->K, d);
  if (timed) cudaEventRecord(ev0, stream);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = L::DYN_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  // programmatic dependent launch: the prologue of this GEMM overlaps the tail of the kernel before it (pdl_wait in the kernel)
  static const bool no_pdl = getenv("NRV_NO_PDL") != nullptr;   // A/B switch
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = no_pdl ? 1 : 2;
  NRV_CUDA(cudaLaunchKernelEx(&cfg, gemm_kernel<BN, PAIR, DUAL, LNX, PATCH>, ta, tb, to, to2, tx, kp));
  if (timed) cudaEventRecord(ev1, stream);
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

static int gemm_dispatch_native(const nrv_gemm_desc* d, cudaStream_t stream, const PatchView* pv = nullptr);

size_t gemm_workspace_bytes(int M, int N, int K, int dtype) {
  if (dtype != NRV_F32) return 0;
  const size_t Kp = (size_t)((K + 3) / 4) * 4;
  return ((size_t)M + (size_t)N) * 3 * Kp * sizeof(float) + 512;
}

int gemm_dispatch(const nrv_gemm_desc* d, cudaStream_t stream) {
  NRV_REQUIRE(d != nullptr, "nrv_gemm: null descriptor");
  if (d->dtype != NRV_F32) return gemm_dispatch_native(d, stream);
  // ---- check mode: split operands into workspace, then one K-major TF32 GEMM over 3*Kp
  NRV_REQUIRE(d->M > 0 && d->N > 0 && d->K > 0, "nrv_gemm: M,N,K must be positive (got %d,%d,%d)",
              d->M, d->N, d->K);
  NRV_REQUIRE(d->a && d->b && d->out, "nrv_gemm: null operand pointer");
  NRV_REQUIRE(d->workspace != nullptr && d->workspace_bytes >= gemm_workspace_bytes(d->M, d->N, d->K, NRV_F32),
              "nrv_gemm: NRV_F32 operands need a workspace of nrv_gemm_workspace_bytes() bytes");
  const int Kp = (d->K + 3) / 4 * 4;
  float* wa = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(d->workspace) + 255) & ~uintptr_t(255));
  float* wb = wa + (size_t)d->M * 3 * Kp;
  dim3 blk(32, 8);
  split3_kernel<<<dim3((Kp + 31) / 32, (d->M + 31) / 32), blk, 0, stream>>>(
      reinterpret_cast<const float*>(d->a), d->lda, d->a_layout == NRV_MN_MAJOR, d->M, d->K, Kp, 0, wa);
  split3_kernel<<<dim3((Kp + 31) / 32, (d->N + 31) / 32), blk, 0, stream>>>(
      reinterpret_cast<const float*>(d->b), d->ldb, d->b_layout == NRV_MN_MAJOR, d->N, d->K, Kp, 1, wb);
  count_launch(2);
  NRV_CUDA(cudaGetLastError());
  nrv_gemm_desc t = *d;
  t.a = wa; t.lda = 3 * Kp; t.a_layout = NRV_K_MAJOR;
  t.b = wb; t.ldb = 3 * Kp; t.b_layout = NRV_K_MAJOR;
  t.K = 3 * Kp;
  return gemm_dispatch_native(&t, stream);
}

static int gemm_dispatch_native(const nrv_gemm_desc* d, cudaStream_t stream, const PatchView* pv) {
  NRV_REQUIRE(d != nullptr, "nrv_gemm: null descriptor");
  NRV_REQUIRE(d->M > 0 && d->N > 0 && d->K > 0, "nrv_gemm: M,N,K must be positive (got %d,%d,%d)",
              d->M, d->N, d->K);
  NRV_REQUIRE(d->a && d->b && d->out, "nrv_gemm: null operand pointer");
  if (pv != nullptr)
    NRV_REQUIRE(d->dtype == NRV_BF16 && d->N > 128 && d->M > BM && !d->force_single_cta && !d->force_bn128 &&
                    d->ln_stats == nullptr && d->stats_out == nullptr && d->residual == nullptr &&
                    (pv->role == 1 ? d->epi == NRV_EPI_STORE : d->epi == NRV_EPI_ATOMIC_F32),
                "nrv_gemm (patch mode): bf16 pair-kernel shapes with a STORE (forward) / ATOMIC_F32 (weight gradient) epilogue only");
  const int tf32 = d->dtype == NRV_F32 ? 1 : 0;
  NRV_REQUIRE(d->dtype == NRV_BF16 || d->dtype == NRV_F32, "nrv_gemm: dtype must be BF16 or F32");
  const int esz = tf32 ? 4 : 2;
  const int kelems = 128 / esz;  // K elements per 128-byte smem row
  NRV_REQUIRE(d->N % 8 == 0, "nrv_gemm: N must be a multiple of 8 (got %d)", d->N);
  NRV_REQUIRE((d->lda * esz) % 16 == 0 && (d->ldb * esz) % 16 == 0,
              "nrv_gemm: lda/ldb must give 16-byte aligned rows");
  NRV_REQUIRE(((uintptr_t)d->a % 16) == 0 && ((uintptr_t)d->b % 16) == 0 && ((uintptr_t)d->out % 16) == 0,
              "nrv_gemm: operand pointers must be 16-byte aligned");
  NRV_REQUIRE(d->epi >= NRV_EPI_STORE && d->epi <= NRV_EPI_MUL, "nrv_gemm: bad epilogue %d", d->epi);
  const bool out_f32 = d->out_dtype == NRV_F32 || d->epi == NRV_EPI_ATOMIC_F32;
  NRV_REQUIRE(d->ldo % (out_f32 ? 4 : 8) == 0, "nrv_gemm: ldo must keep 16-byte aligned rows");
  if (d->epi == NRV_EPI_DGELU || d->epi == NRV_EPI_MUL)
    NRV_REQUIRE(d->aux != nullptr && d->ldaux % 8 == 0, "nrv_gemm: DGELU / MUL need aux");
  if (d->epi == NRV_EPI_GELU_GRAD) NRV_REQUIRE(d->out2 != nullptr, "nrv_gemm: GELU_GRAD needs out2");
  if (d->residual) NRV_REQUIRE(d->ldr % (out_f32 ? 4 : 8) == 0, "nrv_gemm: ldr alignment");

  const int BN = (d->N > 128 && !d->force_bn128) ? 256 : 128;
  // CTA-pair kernel (cta_group::2, 256-row units) for everything that has at least two row tiles
  static const bool env_single = getenv("NRV_GEMM_SINGLE_CTA") != nullptr;   // A/B switch for tuning
  const bool pair = BN == 256 && d->M > BM && !d->force_single_cta && (!env_single || pv != nullptr);
  const int esz0 = tf32 ? 4 : 2;
  const int kb_total0 = (d->K + 128 / esz0 - 1) / (128 / esz0);
  // DUAL (512 x 256 super tile, B shared by two row blocks): where a unit is long enough to amortise the read-out
  // bubble and the row count does not strand more than ~5 % of the MMAs in a dead second row block.
  // tile_mode: 0 auto, 1 never, 2 always (when the pair kernel applies); NRV_GEMM_DUAL=0/1 overrides auto.
  const bool lnx = d->ln_stats != nullptr || d->stats_out != nullptr;   // folded-LayerNorm epilogues: their own instantiation
  bool dual = false;
  int dual_half = 0, dual_rr = 1;
  if (pair && d->tile_mode != 1 && !lnx && pv == nullptr) {
    static const char* env_dual = getenv("NRV_GEMM_DUAL");   // A/B switch for tuning: 0 never, 1 wherever it applies
    // Cost model in units of one 256x256x64 K block (512 MMA cycles), waves over the CTA pairs of the device.
    // Measured (profiles/r2_gemm_dual_tiles.txt, K-major A, K >= 2304): a 512x256 unit costs 0.85-0.9 of two 256x256
    // units, plus the read-out bubble of ~3 K blocks; split-K products (MN-major operands) showed no gain.
    const int pairs = num_sms() / 2;
    const int nt = (d->N + BN - 1) / BN;
    const int rb = (d->M + 2 * BM - 1) / (2 * BM);                       // 256-row blocks
    const long long waves_c = ((long long)rb * nt + pairs - 1) / pairs;
    const double cost_c = (double)waves_c * kb_total0;
    // DUAL: an odd row-block count leaves nt half-dead units (one MMA per K step, the cost of a classic unit); they are
    // dealt to the pairs that carry one unit more than the rest (unit_decode).  Makespan by simulating the round-robin.
    const int units_d = ((rb + 1) / 2) * nt;
    const int nclu_d = units_d < pairs ? units_d : pairs;
    if ((rb & 1) && d->epi != NRV_EPI_ATOMIC_F32 && nt <= 16 && units_d > nclu_d) {
      const int rr = units_d % nclu_d != 0 ? units_d % nclu_d : nclu_d;
      bool ok = true;                                       // every half slot must be a slot of this launch
      for (int h = 0; h < nt; ++h) ok = ok && ((h >> 1) % rr) + ((h & 1) + 2 * ((h >> 1) / rr)) * nclu_d < units_d;
      if (ok) { dual_half = nt; dual_rr = rr; }
    }
    // The decision keeps the plain wave count: with the half units balanced the headline shapes (T = 50432: 297 units
    // on 74 pairs, makespan 4 full units instead of 4.5) should have won by ~10 %, but measured inside the step the dual
    // tiling lost on all three (fc2 fwd 205 -> 256 us, fc1 dX 196 -> 205, qkv dX 152 -> 164; step -1.2 %,
    // profiles/r2c_dual_balance_negative_result.txt), so the balancing only removes the worst case where dual is chosen anyway.
    const long long waves_d = ((long long)units_d + pairs - 1) / pairs;
    const double cost_d = (double)waves_d * (2.0 * 0.87 * kb_total0 + 3.0);
    const bool epi_ok = d->epi == NRV_EPI_STORE && d->pos_rows_in <= 0;
    dual = epi_ok && kb_total0 >= 24 && cost_d < 0.97 * cost_c;
    if (env_dual) dual = env_dual[0] == '1' && (d->epi == NRV_EPI_STORE || d->epi == NRV_EPI_ATOMIC_F32);
    if (d->tile_mode == 2) dual = true;
    static const bool no_balance = getenv("NRV_GEMM_DUAL_NOBALANCE") != nullptr;   // A/B switch
    if (!dual || d->epi == NRV_EPI_ATOMIC_F32 || no_balance) { dual_half = 0; dual_rr = 1; }
  }
  const int bm_unit = pair ? (dual ? 4 * BM : 2 * BM) : BM;

  GemmKernelParams kp{};
  kp.M = d->M; kp.N = d->N; kp.K = d->K;
  kp.num_m_tiles = (d->M + bm_unit - 1) / bm_unit;
  kp.num_n_tiles = (d->N + BN - 1) / BN;
  kp.kb_total = (d->K + kelems - 1) / kelems;
  kp.dual_half = dual_half; kp.dual_rr = dual_rr;
  if (pv != nullptr) {
    kp.patch_role = pv->role;
    kp.pm_npx = pv->npx; kp.pm_nb = pv->nb; kp.pm_nxg = pv->gw / pv->npx; kp.pm_gh = pv->gh; kp.pm_gw = pv->gw;
    kp.pm_kbc = pv->ph * pv->pw / 64; kp.pm_rp1 = 64 / pv->pw; kp.pm_ph = pv->ph; kp.pm_B = pv->B;
    kp.pm_tpi = pv->tokens_per_img; kp.pm_tok_off = pv->tok_off;
    const uint32_t line = (uint32_t)pv->pw * 2;                        // bytes of one patch row = one smem line
    const uint32_t rows = pv->role == 1 ? (uint32_t)BM : 64u;          // lines per sub-tile: 128 patches (A tile) / 64 (a K block of dW)
    kp.pm_layout = (uint32_t)pv->layout; kp.pm_sbo = 8 * line; kp.pm_lbo = rows * line;
    for (int k = 0; k < 4; ++k)
      kp.pm_koff[k] = pv->role == 1 ? (uint32_t)((k * 16) / pv->pw) * rows * line + (uint32_t)((k * 16) % pv->pw) * 2   // K-major: sub-tile, then along the line
                                    : (uint32_t)k * 16 * line;                                                           // MN-major: 16 patch rows further
  }
  kp.a_mn = d->a_layout == NRV_MN_MAJOR;
  kp.b_mn = d->b_layout == NRV_MN_MAJOR;
  kp.tf32 = tf32;

  const int sms = pair ? num_sms() / 2 : num_sms();   // schedulable units per wave (CTAs or CTA pairs)
  int splits = d->splits;
  const int tiles = kp.num_m_tiles * kp.num_n_tiles;
  if (d->epi != NRV_EPI_ATOMIC_F32) {
    splits = 1;
  } else if (splits <= 0) {
    // pick the K-split count that minimises (waves / splits): dW outputs are a handful of tiles
    // with a very long reduction, so the K range is what fills the 148 SMs
    int max_splits = kp.kb_total / 8;
    if (max_splits < 1) max_splits = 1;
    if (max_splits > 64) max_splits = 64;
    float best = 1e30f;
    splits = 1;
    for (int s = 1; s <= max_splits; ++s) {
      const int waves = (tiles * s + sms - 1) / sms;
      const float cost = (float)waves / (float)s + 0.004f * s;  // small penalty: extra red traffic
      if (cost < best - 1e-6f) { best = cost; splits = s; }
    }
  }
  kp.kb_per_split = (kp.kb_total + splits - 1) / splits;
  kp.splits = (kp.kb_total + kp.kb_per_split - 1) / kp.kb_per_split;  // no empty splits

  // smem descriptor geometry (bytes); k-step = one UMMA K (32 bytes of K)
  const uint32_t chunk_bytes = BK_BYTES * kelems;  // one MN-major chunk: kelems k-rows x 128 B
  if (!kp.a_mn) { kp.a_kstep = 32 >> 4; kp.a_lbo = 16; kp.a_sbo = 1024; }
  else          { kp.a_kstep = (8 * 128 * (tf32 ? 1 : 2)) >> 4; kp.a_lbo = chunk_bytes; kp.a_sbo = 1024; }
  if (!kp.b_mn) { kp.b_kstep = 32 >> 4; kp.b_lbo = 16; kp.b_sbo = 1024; }
  else          { kp.b_kstep = (8 * 128 * (tf32 ? 1 : 2)) >> 4; kp.b_lbo = chunk_bytes; kp.b_sbo = 1024; }
  kp.idesc = make_idesc(tf32 ? 2u : 1u, kp.a_mn, kp.b_mn, pair ? 2 * BM : BM, BN);

  kp.epi = d->epi;
  kp.alpha = d->alpha;
  kp.out = d->out; kp.ldo = d->ldo; kp.out2 = d->out2;
  kp.bias = d->bias;
  kp.residual = d->residual; kp.ldr = d->ldr;
  kp.aux = d->aux; kp.ldaux = d->ldaux;
  kp.colsum = d->colsum;
  if (d->colsum != nullptr)
    NRV_REQUIRE(d->epi == NRV_EPI_MUL && !out_f32 && ((uintptr_t)d->colsum % 8) == 0 && d->N % 2 == 0,
                "nrv_gemm: colsum needs EPI_MUL with bf16 output and an 8-byte aligned fp32 vector");
  kp.pos = d->pos; kp.pos_rows_in = d->pos_rows_in; kp.pos_rows_out = d->pos_rows_out;
  kp.pos_row_off = d->pos_row_off; kp.ldpos = d->ldpos;
  kp.out_f32 = out_f32 ? 1 : 0;
  kp.ln_stats = d->ln_stats; kp.ln_eps = d->ln_eps;
  kp.ln_inv_dim = d->ln_stats ? 1.0f / (float)d->K_ln : 0.f;
  kp.ln_mean_out = d->ln_mean_out; kp.ln_rstd_out = d->ln_rstd_out;
  kp.stats_out = d->stats_out;
  if (d->ln_stats) {
    NRV_REQUIRE(d->bias != nullptr && d->K_ln > 0 && d->alpha == 1.0f,
                "nrv_gemm: a folded LayerNorm needs bias (= c), K_ln (the normalised width) and alpha = 1");
    NRV_REQUIRE(((uintptr_t)d->ln_stats % 16) == 0, "nrv_gemm: ln_stats must be 16-byte aligned");
    NRV_REQUIRE((d->ln_mean_out == nullptr) == (d->ln_rstd_out == nullptr), "nrv_gemm: ln_mean_out and ln_rstd_out come together");
    NRV_REQUIRE(d->epi == NRV_EPI_STORE || d->epi == NRV_EPI_GELU || d->epi == NRV_EPI_GELU_GRAD,
                "nrv_gemm: a folded LayerNorm needs a STORE / GELU / GELU_GRAD epilogue");
  }
  if (d->stats_out)
    NRV_REQUIRE(d->epi == NRV_EPI_STORE && d->pos_rows_in <= 0 && ((uintptr_t)d->stats_out % 16) == 0,
                "nrv_gemm: stats_out needs the plain / residual STORE epilogue and a 16-byte aligned buffer");
  if (d->pos) NRV_REQUIRE(d->ldpos % 4 == 0 && (d->pos_rows_in > 0 || pv != nullptr), "nrv_gemm: pos table alignment");

  const CUtensorMapDataType dt = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap ta, tb;
  int rc;
  // patch mode: the image as a 5-D tensor (p2, px, image, y, channel); a box = 64 / pw rows of npx neighbouring patches of
  // nimg images: one sub-tile [patches x pw*2 bytes] per patch row (see PatchView)
  auto image_map = [&](CUtensorMap* m, uint32_t nimg) {
    const uint64_t dims[5] = {(uint64_t)pv->pw, (uint64_t)pv->gw, (uint64_t)pv->B, (uint64_t)pv->H, (uint64_t)pv->C};
    const uint64_t strides[4] = {(uint64_t)pv->pw * 2, (uint64_t)pv->C * pv->H * pv->W * 2, (uint64_t)pv->W * 2,
                                 (uint64_t)pv->H * pv->W * 2};
    const uint32_t box[5] = {(uint32_t)pv->pw, (uint32_t)pv->npx, nimg, (uint32_t)(64 / pv->pw), 1u};
    const CUtensorMapSwizzle swz = pv->layout == 6 ? CU_TENSOR_MAP_SWIZZLE_32B
                                   : (pv->layout == 4 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
    return encode_tmap_5d(m, dt, pv->img, dims, strides, box, swz);
  };
  if (pv != nullptr && pv->role == 1) rc = image_map(&ta, (uint32_t)pv->nb);
  else if (pv != nullptr)   // dW: gradient rows [B * tokens_per_img, lda] as (d, token, image), gathered in the image tiles' row order
    rc = encode_tmap_3d(&ta, dt, d->a, (uint64_t)d->M, (uint64_t)pv->tokens_per_img, (uint64_t)pv->B, (uint64_t)d->lda * esz,
                        (uint64_t)d->lda * esz * pv->tokens_per_img, 64, (uint32_t)pv->npx, (uint32_t)(pv->nb / 2), CU_TENSOR_MAP_SWIZZLE_128B);
  else if (!kp.a_mn) rc = encode_tmap_2d(&ta, dt, d->a, d->K, d->M, (uint64_t)d->lda * esz, kelems, BM, CU_TENSOR_MAP_SWIZZLE_128B);
  else          rc = encode_tmap_2d(&ta, dt, d->a, d->M, d->K, (uint64_t)d->lda * esz, kelems, kelems, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  if (pv != nullptr && pv->role == 2) rc = image_map(&tb, (uint32_t)(pv->nb / 2));
  else if (!kp.b_mn) rc = encode_tmap_2d(&tb, dt, d->b, d->K, d->N, (uint64_t)d->ldb * esz, kelems, pair ? BN / 2 : BN, CU_TENSOR_MAP_SWIZZLE_128B);
  else          rc = encode_tmap_2d(&tb, dt, d->b, d->N, d->K, (uint64_t)d->ldb * esz, kelems, kelems, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;

  // output path: TMA store unless the epilogue needs per-row scatter (token remap) or atomics
  kp.tma_epi = (d->epi != NRV_EPI_ATOMIC_F32 && d->pos_rows_in <= 0) ? 1 : 0;
  if (!kp.tma_epi)
    NRV_REQUIRE(d->epi == NRV_EPI_ATOMIC_F32 || (d->epi == NRV_EPI_STORE && d->residual == nullptr),
                "nrv_gemm: the token-remap epilogue supports EPI_STORE without residual only");
  NRV_REQUIRE(kp.tma_epi || d->epi == NRV_EPI_STORE || d->epi == NRV_EPI_ATOMIC_F32, "nrv_gemm: epilogue %d needs the TMA path", d->epi);
  CUtensorMap to, to2, tx;
  memset(&to, 0, sizeof(to));
  memset(&to2, 0, sizeof(to2));
  memset(&tx, 0, sizeof(tx));
  if (kp.tma_epi) {
    const CUtensorMapDataType odt = out_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const int osz = out_f32 ? 4 : 2;
    const uint32_t bw = out_f32 ? 32 : 64;  // 128 bytes of output columns per row
    if (pv != nullptr) {   // patch forward: (columns, tokens of an image, images); a warp's 32 rows = w tokens x 32 / w images
      NRV_REQUIRE(!out_f32, "nrv_gemm (patch mode): bf16 output only");
      const uint32_t w = (uint32_t)(pv->npx < 32 ? pv->npx : 32);
      rc = encode_tmap_3d(&to, odt, d->out, (uint64_t)d->N, (uint64_t)pv->tokens_per_img, (uint64_t)pv->B, (uint64_t)d->ldo * osz,
                          (uint64_t)d->ldo * osz * pv->tokens_per_img, bw, w, 32 / w, CU_TENSOR_MAP_SWIZZLE_128B);
    } else {
      rc = encode_tmap_2d(&to, odt, d->out, d->N, d->M, (uint64_t)d->ldo * osz, bw, 32, CU_TENSOR_MAP_SWIZZLE_128B);
    }
    if (rc) return rc;
    if ((d->epi == NRV_EPI_GELU || d->epi == NRV_EPI_GELU_GRAD) && d->out2 != nullptr) {
      NRV_REQUIRE(((uintptr_t)d->out2 % 16) == 0, "nrv_gemm: out2 must be 16-byte aligned");
      rc = encode_tmap_2d(&to2, odt, d->out2, d->N, d->M, (uint64_t)d->ldo * osz, bw, 32, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    } else {
      to2 = to;
    }
    // tile-shaped epilogue input (same geometry as the output tiles)
    tx = to;
    const void* xptr = nullptr;
    long long xld = 0;
    if (d->epi == NRV_EPI_DGELU) { xptr = d->aux; xld = d->ldaux; kp.extra = 2; }
    else if (d->epi == NRV_EPI_MUL) { xptr = d->aux; xld = d->ldaux; kp.extra = 3; }
    else if (d->residual != nullptr) { xptr = d->residual; xld = d->ldr; kp.extra = 1; }
    NRV_REQUIRE(!(d->epi != NRV_EPI_STORE && d->residual != nullptr), "nrv_gemm: a residual needs EPI_STORE");
    if (xptr != nullptr) {
      NRV_REQUIRE(((uintptr_t)xptr % 16) == 0 && (xld * osz) % 16 == 0, "nrv_gemm: residual / aux alignment");
      rc = encode_tmap_2d(&tx, odt, xptr, d->N, d->M, (uint64_t)xld * osz, bw, 32, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
  }

  const int units = tiles * kp.splits;
  const int grid = units < sms ? units : sms;
  if (pv != nullptr) {
    NRV_REQUIRE(pair && !dual && !lnx, "nrv_gemm (patch mode): internal dispatch error");
    return launch<256, true, false, false, true>(d, kp, ta, tb, to, to2, tx, 2 * grid, stream);
  }
  if (lnx) {
    if (pair) return launch<256, true, false, true>(d, kp, ta, tb, to, to2, tx, 2 * grid, stream);
    if (BN == 256) return launch<256, false, false, true>(d, kp, ta, tb, to, to2, tx, grid, stream);
    return launch<128, false, false, true>(d, kp, ta, tb, to, to2, tx, grid, stream);
  }
  if (pair && dual) return launch<256, true, true, false>(d, kp, ta, tb, to, to2, tx, 2 * grid, stream);
  if (pair) return launch<256, true, false, false>(d, kp, ta, tb, to, to2, tx, 2 * grid, stream);
  if (BN == 256) return launch<256, false, false, false>(d, kp, ta, tb, to, to2, tx, grid, stream);
  return launch<128, false, false, false>(d, kp, ta, tb, to, to2, tx, grid, stream);
}

int gemm_dispatch_patch(const nrv_gemm_desc* d, const PatchView* pv, cudaStream_t stream) { return gemm_dispatch_native(d, stream, pv); }

bool patch_tma_shape_ok(int C, int H, int W, int ph, int pw, int order, int dtype, int img_dtype, int D) {
  static const bool off = getenv("NRV_NO_PATCH_TMA") != nullptr;   // A/B switch: materialise the patch matrix (im2col kernel)
  if (off || dtype != NRV_BF16 || img_dtype != NRV_BF16 || order != NRV_PATCH_CP1P2) return false;
  if ((pw != 16 && pw != 32 && pw != 64) || ph % (64 / pw) != 0) return false;   // 64 patch elements = whole rows of one patch
  if (ph <= 0 || H % ph != 0 || W % pw != 0 || (W * 2) % 16 != 0) return false;
  if ((C * ph * pw) <= 128 || D <= 128 || D % 8 != 0) return false;             // the CTA-pair kernel's shapes
  if (H > 65535 || W / pw > 65535 || C > 65535) return false;
  return true;
}

int patch_view_init(PatchView* pv, const void* img, int B, int C, int H, int W, int ph, int pw, int tokens_per_img, int tok_off, int role) {
  NRV_REQUIRE(img != nullptr && ((uintptr_t)img % 16) == 0, "patch embedding (TMA): the image must be 16-byte aligned");
  NRV_REQUIRE(role == 1 || role == 2, "patch embedding (TMA): bad role");
  pv->img = img; pv->B = B; pv->C = C; pv->H = H; pv->W = W; pv->ph = ph; pv->pw = pw;
  pv->gh = H / ph; pv->gw = W / pw;
  int npx = 1;
  while (npx < 64 && pv->gw % (2 * npx) == 0) npx *= 2;
  pv->npx = npx; pv->nb = 128 / npx;
  pv->tokens_per_img = tokens_per_img; pv->tok_off = tok_off; pv->role = role;
  pv->layout = pw == 16 ? 6 : (pw == 32 ? 4 : 2);
  return NRV_OK;
}

}  // namespace nrv

using namespace nrv;

extern "C" {

int nrv_patch_embed_supported(int C, int H, int W, int ph, int pw, int patch_order, int dtype, int img_dtype, int D) {
  return patch_tma_shape_ok(C, H, W, ph, pw, patch_order, dtype, img_dtype, D) ? 1 : 0;
}

int nrv_patch_embed_fwd(const void* img, int B, int C, int H, int W, int ph, int pw, const void* w, long long ldw,
                        const float* bias, const float* pos, long long ldpos, int tokens_per_img, int tok_off, void* out,
                        long long ldo, int D, void* stream) {
  int rc = require_init();
  if (rc) return rc;
  NRV_REQUIRE(img && w && out, "nrv_patch_embed_fwd: null pointer");
  NRV_REQUIRE(patch_tma_shape_ok(C, H, W, ph, pw, NRV_PATCH_CP1P2, NRV_BF16, NRV_BF16, D),
              "nrv_patch_embed_fwd: unsupported shape (see nrv_patch_embed_supported); use nrv_im2col + nrv_gemm");
  NRV_REQUIRE(tok_off >= 0 && tokens_per_img >= tok_off + (H / ph) * (W / pw), "nrv_patch_embed_fwd: token rows do not hold the patches");
  PatchView pv;
  rc = patch_view_init(&pv, img, B, C, H, W, ph, pw, tokens_per_img, tok_off, 1);
  if (rc) return rc;
  const int tiles = pv.gh * (pv.gw / pv.npx) * ((B + pv.nb - 1) / pv.nb);
  nrv_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.M = tiles * 128; d.N = D; d.K = C * ph * pw;
  d.dtype = NRV_BF16; d.out_dtype = NRV_BF16;
  d.a = img; d.lda = d.K; d.a_layout = NRV_K_MAJOR;
  d.b = w; d.ldb = ldw; d.b_layout = NRV_K_MAJOR;
  d.epi = NRV_EPI_STORE; d.alpha = 1.0f;
  d.out = out; d.ldo = ldo; d.bias = bias;
  d.pos = pos; d.ldpos = ldpos;
  d.tile_mode = 1;
  if (pos) NRV_REQUIRE(ldpos % 4 == 0 && ((uintptr_t)pos % 16) == 0, "nrv_patch_embed_fwd: pos table alignment");
  return gemm_dispatch_patch(&d, &pv, (cudaStream_t)stream);
}

int nrv_patch_embed_bwd_weight(const void* img, int B, int C, int H, int W, int ph, int pw, const void* dx, long long lddx,
                               int tokens_per_img, int tok_off, float* dw, long long lddw, int D, void* stream) {
  int rc = require_init();
  if (rc) return rc;
  NRV_REQUIRE(img && dx && dw, "nrv_patch_embed_bwd_weight: null pointer");
  NRV_REQUIRE(patch_tma_shape_ok(C, H, W, ph, pw, NRV_PATCH_CP1P2, NRV_BF16, NRV_BF16, D),
              "nrv_patch_embed_bwd_weight: unsupported shape (see nrv_patch_embed_supported)");
  NRV_REQUIRE(tok_off >= 0 && tokens_per_img >= tok_off + (H / ph) * (W / pw), "nrv_patch_embed_bwd_weight: token rows do not hold the patches");
  PatchView pv;
  rc = patch_view_init(&pv, img, B, C, H, W, ph, pw, tokens_per_img, tok_off, 2);
  if (rc) return rc;
  const int tiles = pv.gh * (pv.gw / pv.npx) * ((B + pv.nb - 1) / pv.nb);
  nrv_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.M = D; d.N = C * ph * pw; d.K = tiles * 128;
  d.dtype = NRV_BF16; d.out_dtype = NRV_F32;
  d.a = dx; d.lda = lddx; d.a_layout = NRV_MN_MAJOR;
  d.b = img; d.ldb = d.N; d.b_layout = NRV_MN_MAJOR;
  d.epi = NRV_EPI_ATOMIC_F32; d.alpha = 1.0f;
  d.out = dw; d.ldo = lddw;
  d.tile_mode = 1;
  return gemm_dispatch_patch(&d, &pv, (cudaStream_t)stream);
}

}  // extern "C"
