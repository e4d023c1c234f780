// tcgen05 attention backward, second generation: ONE fused pass per (batch, head) item instead of the two
// recompute passes of the first-generation kernel (dQ pass + dK/dV pass; removed in round 2).  Autograd of
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
//   dK_j += dS^T Q_i     A operand = dS^T, bf16, in tensor memory over dP^T (these small-N products are bound by
//                        shared-memory operand reads, so every A operand that can live in TMEM does)
//   dQ_I += dS  K_j      dS^T is ALSO written to shared memory (row = key, 128 B = 64 queries) and read MN-major
//                        (M = queries); issued once per pair of blocks
// B operands are the TMA-loaded Q / dO / K tiles re-read MN-major, exactly as the forward re-reads V.
// Tensor memory (512 columns): two block buffers {S^T 64 | dP^T 64} so block b+1's scores are computed while
// block b is in the SIMT warps; dV_j 64 ; dK_j 64 ; dQ_0 64 ; dQ_1 64  (N <= 256 tokens).
//
// The kernel is HBM-heavy (reads qkv, dO, O; writes dqkv: ~200 KB per item), so nothing is loaded "per item":
//   * a producer warp streams K_j / V_j tiles (ring of 2) and Q_i / dO_i blocks (ring of 5) by TMA, running
//     ahead of the MMA warp across item boundaries; the block sequence is one continuous pipeline
//   * two helper warps prepare the per-query vectors (-lse*log2e, -delta) of the NEXT item (double buffered)
//   * accumulators leave through swizzled staging (the idle dS pair buffer) and TMA stores; rows beyond N are
//     clipped by the tensor map
// Zero padding does the masking: rows >= N of every tile are zero-filled by TMA and the padded vector entries
// are 0, so padded queries give dP = 0, delta = 0 -> dS = 0 and meet dO = 0; padded keys meet K = 0 in dQ.
#include "common.cuh"
#include "nrvit_internal.h"

namespace nrv {

constexpr int B2_DH = 64;
constexpr int B2_SIMT_WARPS = 8;
constexpr int B2_W_MMA = 8, B2_W_TMA = 9, B2_W_VEC = 10;   // warp roles (two vector-helper warps: 10, 11)
constexpr int B2_THREADS = 32 * 12;
constexpr int B2_CHUNK = 128 * 128;                     // [128 rows x 64 bf16] swizzled tile, bytes
constexpr int B2_KV_SLOT = 2 * B2_CHUNK;                // K_j | V_j
constexpr int B2_QD_HALF = 64 * 128;                    // [64 rows x 64 bf16]
constexpr int B2_QD_SLOT = 2 * B2_QD_HALF;              // Q_i | dO_i
constexpr int B2_QD_SLOTS = 5;                          // MMA1 runs two blocks ahead of MMA2: 3 slots in use + 2 in flight

struct Bwd2Params {
  int B, N, H, NP;       // NP = N rounded up to 16 (<= 256)
  int KT;                // key tiles of 128 rows
  int NQB;               // query blocks of 64 columns (the last one may be 16..64 wide)
  int items;             // B * H
  float scale, scale_log2e;
  const bf16* o;         // [B, N, H*dh]
  const bf16* dout;      // [B, N, H*dh]
  const float* lse;      // [B, H, N]
  long long* dbg;        // optional phase timestamps of CTA 0
  float* dbias;          // optional [3*H*dh] fp32: += column sums of the stored dQ|dK|dV (the in_proj bias gradient)
};

struct Bwd2Smem {
  static constexpr int off_kv = 0;                                   // 2 slots
  static constexpr int off_qd = off_kv + 2 * B2_KV_SLOT;             // 5 slots
  static constexpr int off_ds = off_qd + B2_QD_SLOTS * B2_QD_SLOT;   // 2 pair buffers x 2 chunks
  static constexpr int off_vec = off_ds + 4 * B2_CHUNK;              // [2 items][nlse 256 | ndel 256] floats
  static constexpr int off_bar = off_vec + 2 * 2 * 256 * 4;
  static constexpr int n_bars = 26;
  static constexpr int total = off_bar + n_bars * 8 + 16 + 1024;
};

__device__ __forceinline__ float b2_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(B2_THREADS, 1)
attn_bwd2_kernel(const __grid_constant__ CUtensorMap tm_kv, const __grid_constant__ CUtensorMap tm_q,
                 const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ CUtensorMap tm_out,
                 const Bwd2Params p) {
  pdl_launch_dependents();   // the next kernel in the stream (a PDL-launched GEMM) may start its prologue
  using L = Bwd2Smem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const int NP = p.NP, N = p.N, H = p.H, KT = p.KT, NQB = p.NQB;
  const uint32_t sKV = sbase + L::off_kv, sQD = sbase + L::off_qd, sdS = sbase + L::off_ds;
  const uint32_t bar0 = sbase + L::off_bar;
  // barrier map (8 bytes each)
  const uint32_t kv_full = bar0, kv_empty = bar0 + 16, qd_full = bar0 + 32, qd_empty = bar0 + 72, bar_s0 = bar0 + 112,
                 bar_p0 = bar0 + 128, bar_row = bar0 + 144, bar_accfree = bar0 + 152, vec_full = bar0 + 160,
                 vec_empty = bar0 + 176;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + L::off_bar + L::n_bars * 8);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // the dS pair buffers feed the dQ product with up to 64 stale query columns when a pair is incomplete, and with
  // stale rows where a warp of padded keys skipped its arithmetic: start them finite
  for (int i = threadIdx.x; i < 4 * B2_CHUNK / 16; i += B2_THREADS)
    *reinterpret_cast<uint4*>(smem + L::off_ds + i * 16) = make_uint4(0, 0, 0, 0);
  fence_async_smem();
  if (warp == B2_W_MMA) {
    if (elect_one()) {
      tma_prefetch_desc(&tm_kv);
      tma_prefetch_desc(&tm_q);
      tma_prefetch_desc(&tm_do);
      tma_prefetch_desc(&tm_out);
      for (int s = 0; s < 2; ++s) {
        mbar_init(kv_full + 8 * s, 1);
        mbar_init(kv_empty + 8 * s, 1);
        mbar_init(bar_s0 + 8 * s, 1);
        mbar_init(bar_p0 + 8 * s, B2_SIMT_WARPS);
        mbar_init(vec_full + 8 * s, 64);
        mbar_init(vec_empty + 8 * s, B2_SIMT_WARPS);
      }
      for (int s = 0; s < B2_QD_SLOTS; ++s) {
        mbar_init(qd_full + 8 * s, 1);
        mbar_init(qd_empty + 8 * s, 1);
      }
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
  const int TB = my_items * NB;                 // blocks of this CTA: one continuous pipeline

  if (warp == B2_W_TMA) {
    // ================================ TMA producer (one lane) ===================================
    if (elect_one()) {
      int kvc = 0, qs = 0;
      uint32_t qph = 1;                // parity to wait on qd_empty: flips each time the ring wraps
      for (int li = 0; li < my_items; ++li) {
        const int item = (int)blockIdx.x + li * (int)gridDim.x;
        const int b = item / H, h = item % H;
        for (int j = 0; j < KT; ++j) {
          {
            const int s = kvc & 1;
            mbar_wait(kv_empty + 8 * s, ((kvc >> 1) & 1) ^ 1, 30);
            mbar_arrive_expect_tx(kv_full + 8 * s, 2 * B2_CHUNK);
            tma_load_3d(sKV + s * B2_KV_SLOT, &tm_kv, kv_full + 8 * s, (1 * H + h) * B2_DH, j * 128, b);
            tma_load_3d(sKV + s * B2_KV_SLOT + B2_CHUNK, &tm_kv, kv_full + 8 * s, (2 * H + h) * B2_DH, j * 128, b);
            ++kvc;
          }
          for (int i = 0; i < NQB; ++i) {
            const int s = qs;
            mbar_wait(qd_empty + 8 * s, qph, 31);
            mbar_arrive_expect_tx(qd_full + 8 * s, 2 * B2_QD_HALF);
            tma_load_3d(sQD + s * B2_QD_SLOT, &tm_q, qd_full + 8 * s, (0 * H + h) * B2_DH, i * 64, b);
            tma_load_3d(sQD + s * B2_QD_SLOT + B2_QD_HALF, &tm_do, qd_full + 8 * s, h * B2_DH, i * 64, b);
            if (++qs == B2_QD_SLOTS) { qs = 0; qph ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == B2_W_MMA) {
    // ================================ MMA issue (one lane) ======================================
    if (elect_one() && TB > 0) {
      const uint32_t idesc_acc = make_idesc(1u, 0u, 1u, 128u, 64u);    // dV / dK: A K-major (TMEM or smem), B MN-major
      const uint32_t idesc_dq = make_idesc(1u, 1u, 1u, 128u, 64u);     // dQ: A MN-major, B MN-major
      // descriptors = fixed bits + (address >> 4)
      const uint64_t dfix = make_smem_desc_sw128(0, 16, 1024);
      const uint64_t dfix_mn2 = make_smem_desc_sw128(0, B2_CHUNK, 1024);   // MN-major A over two 64-wide chunks
      auto D = [&](uint32_t addr) { return dfix + (uint64_t)(addr >> 4); };
      int n_lb = 0, n_row = 0, n_i = 0; // decode state for mma1: local block, global row and query block of the next block to issue
      int n_qs = 0;                    // ... and its Q/dO ring slot / parity
      uint32_t n_qph = 0;
      auto mma1 = [&](int g) {         // S^T and dP^T of global block g into TMEM buffer g & 1
        const int i = n_i;              // (running counters: an integer division per block costs the issuing thread ~40 cycles)
        const int ks = n_row & 1, qs = n_qs;
        if (i == 0) mbar_wait(kv_full + 8 * ks, (n_row >> 1) & 1, 11);
        mbar_wait(qd_full + 8 * qs, n_qph, 12);
        if (++n_qs == B2_QD_SLOTS) { n_qs = 0; n_qph ^= 1; }
        tc_fence_after();
        const uint32_t idesc1 = make_idesc(1u, 0u, 0u, 128u, (uint32_t)min(64, NP - 64 * i));
        const uint32_t tb = T + (uint32_t)(g & 1) * 128u;
        const uint64_t ak = D(sKV + ks * B2_KV_SLOT), av = ak + (B2_CHUNK >> 4);
        const uint64_t bq = D(sQD + qs * B2_QD_SLOT), bd = bq + (B2_QD_HALF >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tb, ak + 2 * k, bq + 2 * k, idesc1, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tb + 64, av + 2 * k, bd + 2 * k, idesc1, k > 0);
        umma_commit(bar_s0 + 8 * (g & 1));
        if (i == NQB - 1) { ++n_row; n_i = 0; } else ++n_i;
        if (++n_lb == NB) n_lb = 0;
      };
      mma1(0);
      if (TB > 1) mma1(1);
      int gp = 0, gr = 0, lb = 0, bj = 0, bi = 0;   // running pair / row counters, local block of g = (key tile bj, query block bi)
      int qs = 0;                        // Q/dO ring slot of block g
      for (int g = 0; g < TB; ++g) {
        const int j = bj, i = bi;
        const int u = g & 1, ks = gr & 1;
        const int ks_q = min(64, NP - 64 * i) / 16;
        long long* dbg = (p.dbg != nullptr && blockIdx.x == 0 && g < 40) ? p.dbg + g * 8 : nullptr;
        mbar_wait(bar_p0 + 8 * u, (g >> 1) & 1, 13);        // P^T in TMEM, dS^T in smem
        if (dbg) dbg[0] = clock64();
        // the accumulators of the previous row (or item) have been read by the epilogue
        if (i == 0 && gr > 0) mbar_wait(bar_accfree, (gr - 1) & 1, 14);
        tc_fence_after();
        if (dbg) dbg[1] = clock64();
        const uint32_t tb = T + (uint32_t)u * 128u;
        const uint32_t ds_pair = sdS + (gp & 1) * 2 * B2_CHUNK;
        const uint64_t b_q = D(sQD + qs * B2_QD_SLOT), b_do = b_q + (B2_QD_HALF >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k)                          // dV_j += P^T dO_i   (K = queries of the block)
          if (k < ks_q) umma_bf16_ts(T_DV, tb + k * 16, b_do + k * (2048 >> 4), idesc_acc, (i > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)                          // dK_j += dS^T Q_i
          if (k < ks_q) umma_bf16_ts(T_DK, tb + 64 + k * 16, b_q + k * (2048 >> 4), idesc_acc, (i > 0 || k > 0) ? 1u : 0u);
        umma_commit(qd_empty + 8 * qs);                      // Q_i / dO_i slot reusable when these retire
        if (++qs == B2_QD_SLOTS) qs = 0;
        if ((i & 1) || i == NQB - 1) {                       // dQ_I += dS K_j over the pair's 128 queries (K = 128 keys)
          const uint64_t a_mn = dfix_mn2 + (uint64_t)(ds_pair >> 4);
          const uint64_t b_k = D(sKV + ks * B2_KV_SLOT);
          const uint32_t t_dq = T_DQ + 64 * (i >> 1);
          // K = the tile's real keys only: rows beyond NP of K_j are zero-filled by TMA, their 16-key steps add nothing
          // (N = 197: 5 of 8 steps in the second tile; N = 64: 4 of 8)
          const int ks_k = min(128, NP - 128 * j) / 16;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k < ks_k) umma_bf16(t_dq, a_mn + k * (2048 >> 4), b_k + k * (2048 >> 4), idesc_dq, (j > 0 || k > 0) ? 1u : 0u);
          ++gp;
        }
        if (i == NQB - 1) {
          umma_commit(bar_row);                              // dV_j / dK_j complete (and dQ when j is the last row)
          umma_commit(kv_empty + 8 * ks);
          ++gr;
        }
        if (g + 2 < TB) mma1(g + 2);
        if (dbg) dbg[2] = clock64();
        if (++bi == NQB) { bi = 0; ++bj; }
        if (++lb == NB) { lb = 0; bj = 0; }
      }
    }
    __syncwarp();
  } else if (warp >= B2_W_VEC) {
    // ================================ per-query vectors of the next item ========================
    // nlse[q] = -lse[q] * log2(e) ; ndel[q] = -<dO_q, O_q> ; zero for q >= N.  64 threads, 4 queries each.
    const int t0 = threadIdx.x - B2_W_VEC * 32;
    const long long HD = (long long)H * B2_DH;
    for (int li = 0; li < my_items; ++li) {
      const int item = (int)blockIdx.x + li * (int)gridDim.x;
      const int b = item / H, h = item % H;
      float* vec = reinterpret_cast<float*>(smem + L::off_vec) + (li & 1) * 512;
      mbar_wait(vec_empty + 8 * (li & 1), ((li >> 1) & 1) ^ 1, 40);
#pragma unroll 1
      for (int t = t0; t < 256; t += 64) {
        float nl = 0.f, nd = 0.f;
        if (t < N) {
          const uint4* po = reinterpret_cast<const uint4*>(p.o + ((long long)b * N + t) * HD + (long long)h * B2_DH);
          const uint4* pd = reinterpret_cast<const uint4*>(p.dout + ((long long)b * N + t) * HD + (long long)h * B2_DH);
          uint4 a[8], c[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) { a[k] = __ldg(po + k); c[k] = __ldg(pd + k); }
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint32_t aw[4] = {a[k].x, a[k].y, a[k].z, a[k].w}, cw[4] = {c[k].x, c[k].y, c[k].z, c[k].w};
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
        vec[t] = nl;
        vec[256 + t] = nd;
      }
      mbar_arrive(vec_full + 8 * (li & 1));
    }
  } else {
    // ================================ SIMT warps: P^T, dS^T, epilogues ==========================
    const int q = warp & 3;                               // TMEM lane quarter
    const int hf = warp >> 2;                             // column half of the block / which accumulator to store
    const int r = q * 32 + lane;                          // key row within the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint64_t c2 = f2_pack(p.scale_log2e, p.scale_log2e);
    uint8_t* dS_gen = smem + L::off_ds;
    int g = 0, gp = 0, gr = 0;
    bool pend = false;                                    // a finished row whose accumulators are not stored yet
    int pend_j = 0, pend_b = 0, pend_h = 0;
    // one epilogue = this warp stores 32 rows x 64 columns of one accumulator through a [32 rows x 128 B] swizzled
    // staging tile and one TMA store.  The staging tile is this warp's 32 rows of the dS pair buffer that is idle
    // while the epilogue runs (see epilogue_row).
    // read-out of one accumulator in two steps (load + pack to bf16 ; stage + store), so that a warp with two of them
    // (item's last row: dV_j or dK_j and a dQ half) loads and packs the second one while the TMA store of the first is
    // still reading the staging tile: the tma_store_wait_read between the two stores shrinks and the MMA warp gets its
    // accumulators back before that wait.
    auto load_pack = [&](uint32_t tsrc, float mul, uint32_t (&w)[32], bool last_read) {
      uint32_t v[2][32];
      tmem_ld_32x32(tsrc + lane_addr, v[0]);
      tmem_ld_32x32(tsrc + lane_addr + 32, v[1]);
      tmem_wait_ld();
      if (last_read) {   // every accumulator of this warp is in registers: the MMA warp may start the next row's products
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_accfree);
      }
      const uint64_t m2 = f2_pack(mul, mul);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          float a, c;
          f2_unpack(f2_mul(f2_pack(__uint_as_float(v[hh][2 * k]), __uint_as_float(v[hh][2 * k + 1])), m2), a, c);
          w[hh * 16 + k] = pack_bf16(a, c);
        }
    };
    auto store_packed = [&](const uint32_t (&w)[32], int col, int row0, int b, uint8_t* stg) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
        sts128(smem_u32(stg) + lane * 128 + ((u ^ (lane & 7)) << 4), w[4 * u], w[4 * u + 1], w[4 * u + 2], w[4 * u + 3]);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(&tm_out, smem_u32(stg), col, row0, b);
        tma_store_commit();
      }
      if (p.dbias != nullptr) {
        // in_proj bias gradient: column sums of the tile as stored (rows < N only; the TMA store clips the rest).  Lane l
        // owns the bf16 pair of columns 2l, 2l+1: word (l & 3) of unit (l >> 2) ^ (row & 7) of each 128-byte row.
        const int nvalid = min(32, N - row0);
        uint64_t acc2 = f2_pack(0.f, 0.f);
        const uint32_t base = smem_u32(stg) + (lane & 3) * 4;
        // eight rows per step: the swizzle term (rr & 7) is a compile-time constant, the eight loads are in flight together
        uint32_t swz[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) swz[k] = base + k * 128 + ((((lane >> 2) ^ k)) << 4);
#pragma unroll 1
        for (int r8 = 0; r8 < nvalid; r8 += 8) {
          uint32_t w[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            w[k] = 0u;
            if (r8 + k < nvalid) asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w[k]) : "r"(swz[k] + r8 * 128) : "memory");
          }
          uint64_t t2[4];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            t2[k] = f2_add(f2_pack(__uint_as_float(w[2 * k] << 16), __uint_as_float(w[2 * k] & 0xffff0000u)),
                           f2_pack(__uint_as_float(w[2 * k + 1] << 16), __uint_as_float(w[2 * k + 1] & 0xffff0000u)));
          acc2 = f2_add(acc2, f2_add(f2_add(t2[0], t2[1]), f2_add(t2[2], t2[3])));
        }
        float s0, s1;
        f2_unpack(acc2, s0, s1);
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p.dbias + col + 2 * lane), "f"(s0), "f"(s1) : "memory");
      }
    };
    // The epilogue of a row runs one block late (after the first block of the next row, or of the next item), so
    // the wait for the row's last products and the stores never sit between two blocks of SIMT work.
    auto epilogue_row = [&](int cur_pb) {
      // Staging: the dS pair buffer NOT used by the block just processed.  Its last readers (the dK / dQ products of
      // the finished row's last pair) retired before bar_row; its next writers are SIMT warps two blocks further on,
      // and neither the warp nor its column-half partner (the only writers of these rows) leaves this epilogue before the
      // TMA store has read the staging (pair barrier below).
      uint8_t* stg = dS_gen + ((cur_pb ^ 1) * 2 + hf) * B2_CHUNK + q * 4096;
      mbar_wait(bar_row, gr & 1, 20);
      tc_fence_after();
      const int j = pend_j, pb = pend_b, ph = pend_h;
      // hf 0 -> dV_j , hf 1 -> dK_j (scaled); warps whose rows are all beyond N have nothing to store.
      // last row of its item: hf 0 -> dQ rows 0..127 , hf 1 -> dQ rows 128..255.
      const bool do_kv = j * 128 + q * 32 < N, do_q = j == KT - 1 && hf * 128 + q * 32 < N;
      if (!do_kv && !do_q) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_accfree);
      } else {
        uint32_t w[32];
        if (do_kv) {
          load_pack(hf == 0 ? T_DV : T_DK, hf == 0 ? 1.f : p.scale, w, !do_q);
          store_packed(w, ((hf == 0 ? 2 : 1) * H + ph) * B2_DH, j * 128 + q * 32, pb, stg);
        }
        if (do_q) {
          load_pack(T_DQ + 64 * hf, p.scale, w, true);        // (under the first store's read of the staging tile)
          if (do_kv) {
            if (lane == 0) tma_store_wait_read<0>();
            __syncwarp();
          }
          store_packed(w, ph * B2_DH, hf * 128 + q * 32, pb, stg);
        }
      }
      if (lane == 0) tma_store_wait_read<0>();
      // a warp's staging rows are next written (as dS^T rows of quarter q) by itself and by its column-half partner only
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      ++gr;
      pend = false;
    };
    for (int li = 0; li < my_items; ++li) {
      const int item = (int)blockIdx.x + li * (int)gridDim.x;
      const int b = item / H, h = item % H;
      const float* nlse_s = reinterpret_cast<const float*>(smem + L::off_vec) + (li & 1) * 512;
      const float* ndel_s = nlse_s + 256;
      mbar_wait(vec_full + 8 * (li & 1), (li >> 1) & 1, 22);
      for (int lb = 0, j = 0, i = 0; lb < NB; ++lb, ++g, (++i == NQB ? (i = 0, ++j) : 0)) {
        const int u = g & 1;
        const int nch = min(64, NP - 64 * i) / 16;          // 16-column chunks in this block (1..4)
        const int c_beg = hf == 0 ? 0 : (nch + 1) / 2, c_end = hf == 0 ? (nch + 1) / 2 : nch;
        long long* sdbg = (p.dbg != nullptr && blockIdx.x == 0 && g < 40 && threadIdx.x == 0) ? p.dbg + g * 8 + 4 : nullptr;
        mbar_wait(bar_s0 + 8 * u, (g >> 1) & 1, 21);
        tc_fence_after();
        if (sdbg) sdbg[0] = clock64();
        const uint32_t tb = T + (uint32_t)u * 128u + lane_addr;
        const int cur_pb = gp & 1;
        uint8_t* ds_row = dS_gen + (cur_pb * 2 + (i & 1)) * B2_CHUNK + r * 128;
        // warps whose 32 key rows are all padding skip the arithmetic: their stale P^T / dS^T rows only reach
        // accumulator rows that are never stored, or meet zero K rows in the dQ product
        const bool rows_live = j * 128 + q * 32 < N;
        if (rows_live && c_beg < c_end) {
          uint32_t s[2][16], d[2][16];
#pragma unroll
          for (int cc = 0; cc < 2; ++cc)
            if (c_beg + cc < c_end) {
              tmem_ld_32x16(tb + (c_beg + cc) * 16, s[cc]);
              tmem_ld_32x16(tb + 64 + (c_beg + cc) * 16, d[cc]);
            }
          tmem_wait_ld();
#pragma unroll
          for (int cc = 0; cc < 2; ++cc)
            if (c_beg + cc < c_end) {
              const int c = c_beg + cc;
              const uint32_t nl4 = smem_u32(nlse_s + i * 64 + c * 16), nd4 = smem_u32(ndel_s + i * 64 + c * 16);
              uint32_t pk[8], dk[8];
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                const float4 l = lds128f(nl4 + 16 * k4), dl = lds128f(nd4 + 16 * k4);
                const float lv[4] = {l.x, l.y, l.z, l.w}, dv[4] = {dl.x, dl.y, dl.z, dl.w};
#pragma unroll
                for (int e = 0; e < 4; e += 2) {
                  float x0, x1, t0v, t1v;
                  f2_unpack(f2_fma(f2_pack(__uint_as_float(s[cc][4 * k4 + e]), __uint_as_float(s[cc][4 * k4 + e + 1])), c2,
                                   f2_pack(lv[e], lv[e + 1])), x0, x1);
                  const float p0 = b2_ex2(x0), p1 = b2_ex2(x1);
                  f2_unpack(f2_mul(f2_pack(p0, p1), f2_add(f2_pack(__uint_as_float(d[cc][4 * k4 + e]), __uint_as_float(d[cc][4 * k4 + e + 1])),
                                                            f2_pack(dv[e], dv[e + 1]))), t0v, t1v);
                  pk[2 * k4 + (e >> 1)] = pack_bf16(p0, p1);
                  dk[2 * k4 + (e >> 1)] = pack_bf16(t0v, t1v);
                }
              }
              // P^T chunk c (bf16 pairs) over the first half of ITS OWN S^T chunk: the other column-half warp
              // never reads these columns
              tmem_st_32x8(tb + c * 16, pk);
              tmem_st_32x8(tb + 64 + c * 16, dk);            // dS^T likewise over its dP^T chunk: A operand of the dK product
              sts128(smem_u32(ds_row) + (((2 * c) ^ (r & 7)) << 4), dk[0], dk[1], dk[2], dk[3]);
              sts128(smem_u32(ds_row) + (((2 * c + 1) ^ (r & 7)) << 4), dk[4], dk[5], dk[6], dk[7]);
            }
          tmem_wait_st();
          fence_async_smem();       // generic-proxy smem writes -> visible to the UMMA operand reads
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p0 + 8 * u);
        if (sdbg) sdbg[1] = clock64();
        if ((i & 1) || i == NQB - 1) ++gp;
        // last block of the item: the vectors are free for the item after next
        if (lb == NB - 1 && lane == 0) mbar_arrive(vec_empty + 8 * (li & 1));
        if (pend) epilogue_row(cur_pb);
        if (i == NQB - 1) { pend = true; pend_j = j; pend_b = b; pend_h = h; }
        if (sdbg) sdbg[2] = clock64();
      }
    }
    if (pend) epilogue_row(gp & 1);   // (gp already points past the last pair: its buffer is the idle one's partner)
    if (lane == 0) tma_store_wait<0>();   // smem must outlive the last bulk store
  }

  tc_fence_before();
  __syncthreads();
  if (warp == B2_W_MMA) tmem_dealloc(T, 512);
}

bool attn_bwd2_supported(int N, int dh, int dtype) { return dtype == NRV_BF16 && dh == B2_DH && N >= 1 && N <= 256; }

int attn_bwd_tc2(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                 int B, int N, int H, int dh, float scale, cudaStream_t st, float* dbias) {
  NRV_REQUIRE(((uintptr_t)dbias % 8) == 0, "tcgen05 attention backward: dbias must be 8-byte aligned");
  NRV_REQUIRE(attn_bwd2_supported(N, dh, NRV_BF16), "tcgen05 attention backward: unsupported shape N=%d dh=%d", N, dh);
  NRV_REQUIRE(((uintptr_t)qkv % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)dout % 16) == 0 && ((uintptr_t)dqkv % 16) == 0,
              "tcgen05 attention: 16-byte alignment");
  Bwd2Params p{};
  p.B = B; p.N = N; p.H = H; p.NP = (N + 15) / 16 * 16;
  p.KT = (N + 127) / 128; p.NQB = (p.NP + 63) / 64; p.items = B * H;
  p.scale = scale; p.scale_log2e = scale * 1.4426950408889634f;
  p.o = (const bf16*)out; p.dout = (const bf16*)dout; p.lse = lse;
  p.dbg = attn_tc_get_debug();
  p.dbias = dbias;
  const uint64_t row_qkv = (uint64_t)3 * H * B2_DH, row_o = (uint64_t)H * B2_DH;
  const CUtensorMapDataType bf = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap tkv, tq, td, tout;
  int rc = encode_tmap_3d(&tkv, bf, qkv, row_qkv, N, B, row_qkv * 2, row_qkv * 2 * N, 64, 128, 1, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = encode_tmap_3d(&tq, bf, qkv, row_qkv, N, B, row_qkv * 2, row_qkv * 2 * N, 64, 64, 1, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = encode_tmap_3d(&td, bf, dout, row_o, N, B, row_o * 2, row_o * 2 * N, 64, 64, 1, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = encode_tmap_3d(&tout, bf, dqkv, row_qkv, N, B, row_qkv * 2, row_qkv * 2 * N, 64, 32, 1, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  const int smem = Bwd2Smem::total;
  static_assert(Bwd2Smem::total <= 227 * 1024, "attention backward: shared memory budget");
  NRV_CUDA(cudaFuncSetAttribute(attn_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = p.items < num_sms() ? p.items : num_sms();
  attn_bwd2_kernel<<<grid, B2_THREADS, smem, st>>>(tkv, tq, td, tout, p);
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

}  // namespace nrv
