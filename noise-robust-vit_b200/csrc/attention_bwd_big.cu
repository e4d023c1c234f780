// tcgen05 attention backward for the shapes attention_bwd2.cu does not cover: head dims 16..80 (multiples of 16) and
// up to 1024 tokens -- ViT-H/14 training (dh = 80, N = 257; vit.py:493-519) ran its attention backward on CUDA cores
// before (307 ms / step at B = 32, almost all of it there).  Autograd of
//   dots = q k^T * scale ; attn = softmax(dots) ; out = attn v          (simple_vit.py:70-75 ; utils.py:207-232)
// with P recomputed from the stored log-sum-exp:
//   P = exp(S*scale - lse) ; dP = dO V^T ; dS = P o (dP - delta) ; delta_q = <dO_q, O_q>
//   dV = P^T dO ; dK = scale * dS^T Q ; dQ = scale * dS K
//
// Same algebra and the same building blocks as attention_bwd2.cu (transposed blocks: key rows in the TMEM lanes, P^T / dS^T
// written back to tensor memory as bf16 and consumed as the A operand of TS-form MMAs), laid out for generality instead of
// overlap: one pipeline per SM, single-buffered, 4 SIMT warps (thread = key row) + 1 control warp (TMA + MMA issue).
//   * a head is ceil(dh/64) column chunks of 64: one 4-D tensor map over [B, N, heads, dh]; columns beyond dh are out of
//     bounds in the innermost dimension and rows beyond N in the token dimension, so TMA zero-fills both and every chunk is a
//     regular [128 rows x 128 B] SWIZZLE_128B tile (as in attention_fwd_big.cu)
//   * blocks = key tile j (128 rows) x query tile I (128 columns):  S^T = K_j Q_I^T and dP^T = V_j dO_I^T (dh/16 K steps)
//     -> SIMT -> dV_j += P^T dO_I, dK_j += dS^T Q_I (TS form, per chunk an N = 64 or N = dh - 64 product) and the block's
//     dQ contribution dS_I K_j (dS^T also goes to shared memory and is read MN-major)
//   * tensor memory: S^T 128 | dP^T 128 | dV_j dh | dK_j dh | dQ dh columns (256 + 3 dh <= 512: dh <= 80)
//   * dQ_I is summed over the key tiles in a per-CTA fp32 scratch (L2-resident: gridDim x N x dh floats): thread = query
//     row, the same thread adds tile after tile, so there are no atomics and the result is run-to-run reproducible; the
//     last key tile's pass scales, rounds and stores it
// Zero padding does the masking exactly as in attention_bwd2.cu; in addition rows of padded keys are written as zeros
// (their K rows are zero too, but exp2(-lse) of a padded key is unbounded).
#include "common.cuh"
#include "nrvit_internal.h"

namespace nrv {

constexpr int BB_THREADS = 32 * 5;
constexpr int BB_TILE = 128 * 128;                   // [128 rows x 64 bf16] swizzled tile, bytes
constexpr uint32_t BB_T_S = 0, BB_T_DP = 128, BB_T_ACC = 256;
constexpr int BB_MAX_N = 1024;

struct BwdBigParams {
  int B, N, H, dh, NP, CH, KT, items;
  float scale, scale_log2e;
  const bf16* o;         // [B, N, H*dh]
  const bf16* dout;      // [B, N, H*dh]
  const float* lse;      // [B, H, N]
  bf16* dqkv;            // [B, N, 3, H, dh]
  float* scratch;        // [gridDim][KT*128][dh]
  AttnDrop dr;           // dropout on the probabilities (dr.p = 0: none): dV = (P o M)^T dO, dP = (dO V^T) o M
  uint32_t dthresh;      // p * 2^24
  float dscale;          // 1 / (1 - p)
};

template <bool DROP>   // DROP: dropout on the probabilities (the mask arithmetic is compiled out of the plain instantiation)
__global__ void __launch_bounds__(BB_THREADS, 1)
attn_bwd_big_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do, const BwdBigParams p) {
  pdl_launch_dependents();   // the next kernel in the stream (a PDL-launched GEMM) may start its prologue
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const int N = p.N, H = p.H, dh = p.dh, CH = p.CH, KT = p.KT, NP = p.NP;
  const int NPQ = KT * 128;
  const uint32_t sK = sbase, sV = sK + CH * BB_TILE, sQ = sV + CH * BB_TILE, sdO = sQ + CH * BB_TILE, sdS = sdO + CH * BB_TILE;
  const int off_ds = 4 * CH * BB_TILE, off_vec = off_ds + 2 * BB_TILE, off_bar = off_vec + 2 * NPQ * 4;
  float* vec = reinterpret_cast<float*>(smem + off_vec);        // nlse[NPQ] | ndel[NPQ]
  const uint32_t bar0 = sbase + off_bar;
  const uint32_t bar_kv = bar0, bar_qd = bar0 + 8, bar_s = bar0 + 16, bar_p = bar0 + 24, bar_o = bar0 + 32;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + off_bar + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // the dS^T tiles feed the dQ product with all 128 key rows and 128 query columns of a block, whether or not every one
  // of them was written for this block: start them finite (stale values of earlier blocks are finite by construction)
  for (int i = threadIdx.x; i < 2 * BB_TILE / 16; i += BB_THREADS)
    *reinterpret_cast<uint4*>(smem + off_ds + i * 16) = make_uint4(0, 0, 0, 0);
  fence_async_smem();
  if (warp == 4) {
    if (elect_one()) {
      tma_prefetch_desc(&tm_qkv);
      tma_prefetch_desc(&tm_do);
      mbar_init(bar_kv, 1);
      mbar_init(bar_qd, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_p, 4);
      mbar_init(bar_o, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t T = *tmem_ptr_smem;
  const uint32_t T_DV = T + BB_T_ACC, T_DK = T_DV + (uint32_t)dh, T_DQ = T_DK + (uint32_t)dh;
  const int my_items = ((int)blockIdx.x < p.items) ? (p.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int ksteps_d = dh / 16;

  if (warp == 4) {
    // ================================ TMA + MMA issue (one lane) ================================
    if (elect_one()) {
      const uint64_t dfix = make_smem_desc_sw128(0, 16, 1024);
      const uint64_t dfix_mn2 = make_smem_desc_sw128(0, BB_TILE, 1024);     // MN-major A over the two 64-query chunks of dS^T
      auto D = [&](uint32_t addr) { return dfix + (uint64_t)(addr >> 4); };
      int g = 0, rowc = 0;
      for (int li = 0; li < my_items; ++li) {
        const int item = (int)blockIdx.x + li * (int)gridDim.x;
        const int b = item / H, h = item % H;
        for (int j = 0; j < KT; ++j) {
          // every reader of the previous K_j / V_j (the products of the previous block) has retired: bar_o was waited
          mbar_arrive_expect_tx(bar_kv, 2 * CH * BB_TILE);
          for (int c = 0; c < CH; ++c) {
            tma_load_4d(sK + c * BB_TILE, &tm_qkv, bar_kv, c * 64, 1 * H + h, j * 128, b);
            tma_load_4d(sV + c * BB_TILE, &tm_qkv, bar_kv, c * 64, 2 * H + h, j * 128, b);
          }
          for (int I = 0; I < KT; ++I, ++g) {
            const int nq = min(128, NP - 128 * I);                  // query columns of this block (multiple of 16)
            const uint32_t ph = (uint32_t)(g & 1);
            mbar_arrive_expect_tx(bar_qd, 2 * CH * BB_TILE);
            for (int c = 0; c < CH; ++c) {
              tma_load_4d(sQ + c * BB_TILE, &tm_qkv, bar_qd, c * 64, 0 * H + h, I * 128, b);
              tma_load_4d(sdO + c * BB_TILE, &tm_do, bar_qd, c * 64, h, I * 128, b);
            }
            if (I == 0) { mbar_wait(bar_kv, (uint32_t)(rowc & 1), 11); ++rowc; }
            mbar_wait(bar_qd, ph, 12);
            tc_fence_after();
            // ---- S^T = K_j Q_I^T ; dP^T = V_j dO_I^T
            const uint32_t idesc1 = make_idesc(1u, 0u, 0u, 128u, (uint32_t)nq);
            for (int ks = 0; ks < ksteps_d; ++ks) {
              const uint32_t o = (uint32_t)((ks >> 2) * BB_TILE + (ks & 3) * 32);
              umma_bf16(T + BB_T_S, D(sK + o), D(sQ + o), idesc1, ks > 0);
            }
            for (int ks = 0; ks < ksteps_d; ++ks) {
              const uint32_t o = (uint32_t)((ks >> 2) * BB_TILE + (ks & 3) * 32);
              umma_bf16(T + BB_T_DP, D(sV + o), D(sdO + o), idesc1, ks > 0);
            }
            umma_commit(bar_s);
            // P^T / dS^T written; the SIMT warps have also read out the previous block's accumulators (program order)
            mbar_wait(bar_p, ph, 13);
            tc_fence_after();
            const int kq = nq / 16;
            for (int c = 0; c < CH; ++c) {
              const uint32_t nc = (uint32_t)min(64, dh - 64 * c);
              const uint32_t idesc_acc = make_idesc(1u, 0u, 1u, 128u, nc);     // A from tensor memory, B MN-major
              const uint32_t idesc_dq = make_idesc(1u, 1u, 1u, 128u, nc);      // A MN-major (dS^T in smem), B MN-major
              for (int k = 0; k < kq; ++k)                             // dV_j += P^T dO_I   (K = queries)
                umma_bf16_ts(T_DV + 64 * c, T + BB_T_S + k * 16, D(sdO + c * BB_TILE + k * 2048), idesc_acc, (I > 0 || k > 0) ? 1u : 0u);
              for (int k = 0; k < kq; ++k)                             // dK_j += dS^T Q_I
                umma_bf16_ts(T_DK + 64 * c, T + BB_T_DP + k * 16, D(sQ + c * BB_TILE + k * 2048), idesc_acc, (I > 0 || k > 0) ? 1u : 0u);
              for (int k = 0; k < 8; ++k)                              // dQ_I (this key tile's share) = dS K_j   (K = 128 keys)
                umma_bf16(T_DQ + 64 * c, dfix_mn2 + (uint64_t)((sdS + k * 2048) >> 4), D(sK + c * BB_TILE + k * 2048), idesc_dq, k > 0 ? 1u : 0u);
            }
            umma_commit(bar_o);
            mbar_wait(bar_o, ph, 14);          // operand tiles and the S^T / dP^T columns are free again
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ================================ SIMT warps: vectors, P^T / dS^T, read-out ==================
    const int q = warp & 3;
    const int r = q * 32 + lane;                           // key row within the tile == TMEM lane; query row in the read-outs
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint64_t c2 = f2_pack(p.scale_log2e, p.scale_log2e);
    const long long HD = (long long)H * dh;
    float* nlse_s = vec;
    float* ndel_s = vec + NPQ;
    int g = 0;
    for (int li = 0; li < my_items; ++li) {
      const int item = (int)blockIdx.x + li * (int)gridDim.x;
      const int b = item / H, h = item % H;
      // ---- per-query vectors of the item: nlse[q] = -lse[q] log2(e) ; ndel[q] = -<dO_q, O_q> ; zero beyond N
      asm volatile("bar.sync 1, 128;" ::: "memory");       // every warp is done with the previous item's vectors
      for (int t = (int)threadIdx.x; t < NPQ; t += 128) {
        float nl = 0.f, nd = 0.f;
        if (t < N) {
          const uint4* po = reinterpret_cast<const uint4*>(p.o + ((long long)b * N + t) * HD + (long long)h * dh);
          const uint4* pd = reinterpret_cast<const uint4*>(p.dout + ((long long)b * N + t) * HD + (long long)h * dh);
          float acc = 0.f;
          for (int k = 0; k < dh / 8; ++k) {
            const uint4 a = __ldg(po + k), c = __ldg(pd + k);
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, cw[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 x = unpack_bf16(aw[e]), y = unpack_bf16(cw[e]);
              acc = fmaf(x.x, y.x, acc);
              acc = fmaf(x.y, y.y, acc);
            }
          }
          nd = -acc;
          nl = -p.lse[((long long)b * H + h) * N + t] * 1.4426950408889634f;
        }
        nlse_s[t] = nl;
        ndel_s[t] = nd;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      float* scr = p.scratch + (size_t)blockIdx.x * NPQ * dh;
      for (int j = 0; j < KT; ++j) {
        const int kidx = j * 128 + r;
        const bool warp_live = j * 128 + q * 32 < N;       // at least one real key in this warp's rows
        const bool row_dead = kidx >= N;
        for (int I = 0; I < KT; ++I, ++g) {
          const int nq = min(128, NP - 128 * I);
          const uint32_t ph = (uint32_t)(g & 1);
          mbar_wait(bar_s, ph, 21);
          tc_fence_after();
          if (warp_live) {
            const uint32_t tS = T + BB_T_S + lane_addr, tP = T + BB_T_DP + lane_addr;
            for (int c0 = 0; c0 < nq / 16; c0 += 2) {
              uint32_t s[2][16], d[2][16];
#pragma unroll
              for (int cc = 0; cc < 2; ++cc)
                if ((c0 + cc) * 16 < nq) {
                  tmem_ld_32x16(tS + (c0 + cc) * 16, s[cc]);
                  tmem_ld_32x16(tP + (c0 + cc) * 16, d[cc]);
                }
              tmem_wait_ld();
#pragma unroll
              for (int cc = 0; cc < 2; ++cc)
                if ((c0 + cc) * 16 < nq) {
                  const int c = c0 + cc;
                  const uint32_t nl4 = smem_u32(nlse_s + I * 128 + c * 16), nd4 = smem_u32(ndel_s + I * 128 + c * 16);
                  uint32_t pk[8], dk[8];
                  // dropout: keep bit t = (query I*128 + 16c + t, this thread's key).  Consecutive queries are N elements
                  // apart in the mask stream, so every element is its own Philox call here (the forward shares one call
                  // among four keys)
                  uint32_t keep = 0xffffu;
                  if constexpr (DROP) {
                    keep = 0u;
                    const unsigned long long e0 = ((((unsigned long long)b * H + h) * N + (unsigned long long)(I * 128 + c * 16)) * N) + (unsigned long long)kidx;
#pragma unroll 4
                    for (int t = 0; t < 16; ++t) {
                      const unsigned long long e = e0 + (unsigned long long)t * N;
                      keep |= ((attn_keep4(p.dr, p.dthresh, e >> 2) >> ((uint32_t)e & 3u)) & 1u) << t;
                    }
                  }
                  const float dsc = DROP ? p.dscale : 1.f;
#pragma unroll
                  for (int k4 = 0; k4 < 4; ++k4) {
                    const float4 l = lds128f(nl4 + 16 * k4), dl = lds128f(nd4 + 16 * k4);
                    const float lv[4] = {l.x, l.y, l.z, l.w}, dv[4] = {dl.x, dl.y, dl.z, dl.w};
#pragma unroll
                    for (int e = 0; e < 4; e += 2) {
                      float x0, x1, t0v, t1v;
                      f2_unpack(f2_fma(f2_pack(__uint_as_float(s[cc][4 * k4 + e]), __uint_as_float(s[cc][4 * k4 + e + 1])), c2,
                                       f2_pack(lv[e], lv[e + 1])), x0, x1);
                      const float p0 = ex2f(x0), p1 = ex2f(x1);
                      if constexpr (DROP) {
                        // mask factors of the two queries: dP o M, and P o M for the dV product
                        const float f0 = ((keep >> (4 * k4 + e)) & 1u) ? dsc : 0.f, f1 = ((keep >> (4 * k4 + e + 1)) & 1u) ? dsc : 0.f;
                        const uint64_t f2 = f2_pack(f0, f1);
                        f2_unpack(f2_mul(f2_pack(p0, p1), f2_fma(f2_pack(__uint_as_float(d[cc][4 * k4 + e]), __uint_as_float(d[cc][4 * k4 + e + 1])), f2,
                                                                  f2_pack(dv[e], dv[e + 1]))), t0v, t1v);
                        pk[2 * k4 + (e >> 1)] = pack_bf16(p0 * f0, p1 * f1);
                      } else {
                        f2_unpack(f2_mul(f2_pack(p0, p1), f2_add(f2_pack(__uint_as_float(d[cc][4 * k4 + e]), __uint_as_float(d[cc][4 * k4 + e + 1])),
                                                                  f2_pack(dv[e], dv[e + 1]))), t0v, t1v);
                        pk[2 * k4 + (e >> 1)] = pack_bf16(p0, p1);
                      }
                      dk[2 * k4 + (e >> 1)] = pack_bf16(t0v, t1v);
                    }
                  }
                  if (row_dead) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) { pk[e] = 0u; dk[e] = 0u; }
                  }
                  // P^T chunk c (bf16 pairs, 8 columns) over the first half of its own S^T chunk, dS^T likewise over dP^T
                  tmem_st_32x8(tS + c * 16, pk);
                  tmem_st_32x8(tP + c * 16, dk);
                  // dS^T to shared memory: row = key, chunk (c >> 2) of 64 queries, 32 bytes = 16 queries per 16-column chunk
                  const uint32_t ds_row = sdS + (uint32_t)((c >> 2) * BB_TILE + r * 128);
                  sts128(ds_row + ((((2 * (c & 3))) ^ (r & 7)) << 4), dk[0], dk[1], dk[2], dk[3]);
                  sts128(ds_row + ((((2 * (c & 3) + 1)) ^ (r & 7)) << 4), dk[4], dk[5], dk[6], dk[7]);
                }
            }
            tmem_wait_st();
            fence_async_smem();       // generic-proxy smem writes -> visible to the UMMA operand reads
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_p);

          mbar_wait(bar_o, ph, 22);
          tc_fence_after();
          // ---- dQ of query tile I: add this key tile's share (thread = query row; the same thread every time)
          if (I * 128 + q * 32 < N) {
            const int qi = I * 128 + r;
            float* sc = scr + (size_t)qi * dh;
            bf16* dst = p.dqkv + (((long long)b * N + qi) * 3 * H + h) * dh;
            for (int c0 = 0; c0 < dh; c0 += 16) {
              uint32_t v[16];
              tmem_ld_32x16(T_DQ + lane_addr + c0, v);
              tmem_wait_ld();
              if (qi < N) {
                float x[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) x[e] = __uint_as_float(v[e]);
                if (j > 0) {
#pragma unroll
                  for (int e = 0; e < 16; e += 4) {
                    const float4 a = *reinterpret_cast<const float4*>(sc + c0 + e);
                    x[e] += a.x; x[e + 1] += a.y; x[e + 2] += a.z; x[e + 3] += a.w;
                  }
                }
                if (j == KT - 1) {
                  uint32_t w[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) w[e] = pack_bf16(x[2 * e] * p.scale, x[2 * e + 1] * p.scale);
                  *reinterpret_cast<uint4*>(dst + c0) = make_uint4(w[0], w[1], w[2], w[3]);
                  *reinterpret_cast<uint4*>(dst + c0 + 8) = make_uint4(w[4], w[5], w[6], w[7]);
                } else {
#pragma unroll
                  for (int e = 0; e < 16; e += 4) *reinterpret_cast<float4*>(sc + c0 + e) = make_float4(x[e], x[e + 1], x[e + 2], x[e + 3]);
                }
              }
            }
          }
          // ---- dV_j / dK_j are complete after the last query tile (thread = key row)
          if (I == KT - 1 && warp_live) {
            bf16* dstv = p.dqkv + (((long long)b * N + kidx) * 3 * H + 2 * H + h) * dh;
            bf16* dstk = p.dqkv + (((long long)b * N + kidx) * 3 * H + 1 * H + h) * dh;
            for (int c0 = 0; c0 < dh; c0 += 16) {
              uint32_t v[16], u[16];
              tmem_ld_32x16(T_DV + lane_addr + c0, v);
              tmem_ld_32x16(T_DK + lane_addr + c0, u);
              tmem_wait_ld();
              if (!row_dead) {
                uint32_t w[8], z[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  w[e] = pack_bf16(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
                  z[e] = pack_bf16(__uint_as_float(u[2 * e]) * p.scale, __uint_as_float(u[2 * e + 1]) * p.scale);
                }
                *reinterpret_cast<uint4*>(dstv + c0) = make_uint4(w[0], w[1], w[2], w[3]);
                *reinterpret_cast<uint4*>(dstv + c0 + 8) = make_uint4(w[4], w[5], w[6], w[7]);
                *reinterpret_cast<uint4*>(dstk + c0) = make_uint4(z[0], z[1], z[2], z[3]);
                *reinterpret_cast<uint4*>(dstk + c0 + 8) = make_uint4(z[4], z[5], z[6], z[7]);
              }
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(T, 512);
}

static int bwd_big_smem(int N, int dh) {
  const int CH = (dh + 63) / 64, KT = (N + 127) / 128;
  return 4 * CH * BB_TILE + 2 * BB_TILE + 2 * KT * 128 * 4 + 128 + 1024;
}

bool attn_bwd_big_supported(int N, int dh, int dtype) {
  if (dtype != NRV_BF16 || dh % 16 != 0 || dh < 16 || dh > 80 || N < 1 || N > BB_MAX_N) return false;
  return bwd_big_smem(N, dh) <= 227 * 1024;
}

size_t attn_bwd_big_scratch_bytes(int B, int N, int H, int dh) {
  const long long items = (long long)B * H;
  const long long grid = items < num_sms() ? items : num_sms();
  return (size_t)grid * (size_t)((N + 127) / 128 * 128) * dh * sizeof(float) + 256;
}

int attn_bwd_big(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* scratch,
                 size_t scratch_bytes, int B, int N, int H, int dh, float scale, cudaStream_t st, float p_drop,
                 unsigned long long seed, int layer) {
  NRV_REQUIRE(attn_bwd_big_supported(N, dh, NRV_BF16), "tcgen05 attention backward (general): unsupported shape N=%d dh=%d", N, dh);
  NRV_REQUIRE(((uintptr_t)qkv % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)dout % 16) == 0 && ((uintptr_t)dqkv % 16) == 0,
              "tcgen05 attention: 16-byte alignment");
  NRV_REQUIRE(scratch != nullptr && scratch_bytes >= attn_bwd_big_scratch_bytes(B, N, H, dh) && ((uintptr_t)scratch % 16) == 0,
              "tcgen05 attention backward (general): workspace of nrv_attn_bwd_workspace() bytes required");
  BwdBigParams p{};
  p.B = B; p.N = N; p.H = H; p.dh = dh; p.NP = (N + 15) / 16 * 16; p.CH = (dh + 63) / 64;
  p.KT = (N + 127) / 128; p.items = B * H;
  p.scale = scale; p.scale_log2e = scale * 1.4426950408889634f;
  p.o = (const bf16*)out; p.dout = (const bf16*)dout; p.lse = lse; p.dqkv = (bf16*)dqkv; p.scratch = scratch;
  p.dr = attn_make_drop(p_drop, seed, layer);
  p.dthresh = (uint32_t)(p_drop * 16777216.0f);
  p.dscale = 1.f / (1.f - p_drop);
  const uint64_t dims[4] = {(uint64_t)dh, (uint64_t)3 * H, (uint64_t)N, (uint64_t)B};
  const uint64_t strides[3] = {(uint64_t)dh * 2, (uint64_t)3 * H * dh * 2, (uint64_t)N * 3 * H * dh * 2};
  const uint64_t dims_o[4] = {(uint64_t)dh, (uint64_t)H, (uint64_t)N, (uint64_t)B};
  const uint64_t strides_o[3] = {(uint64_t)dh * 2, (uint64_t)H * dh * 2, (uint64_t)N * H * dh * 2};
  const uint32_t box[4] = {64, 1, 128, 1};
  CUtensorMap tq, td;
  int rc = encode_tmap_4d(&tq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, qkv, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = encode_tmap_4d(&td, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dout, dims_o, strides_o, box, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  const int smem = bwd_big_smem(N, dh);
  if (p_drop > 0.f) {
    NRV_CUDA(cudaFuncSetAttribute(attn_bwd_big_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  } else {
    NRV_CUDA(cudaFuncSetAttribute(attn_bwd_big_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  const int grid = p.items < num_sms() ? p.items : num_sms();
  if (p_drop > 0.f) attn_bwd_big_kernel<true><<<grid, BB_THREADS, smem, st>>>(tq, td, p);
  else attn_bwd_big_kernel<false><<<grid, BB_THREADS, smem, st>>>(tq, td, p);
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

}  // namespace nrv
