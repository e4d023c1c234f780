// robust=True attention: softmax followed by the reference's 3-iteration Sinkhorn normalisation
//   Q = softmax(Q, -1); 3 x { Q /= sum(Q, -1); Q /= sum(Q, -2) }; Q /= sum(Q, -1)     (utils.py:1031-1037)
// selected by `robust` in both hot-path constructors (simple_vit.py:56-57 ; vit.py:98-110), forward
// and backward, fp32 arithmetic on CUDA cores (both activation dtypes).
//
// Column sums couple every query of a head, so ONE CTA owns the whole N x N probability matrix of one (batch, head):
// in shared memory while it fits (N <= ~204 at dh = 64: 169 KB fp32), otherwise in a per-CTA slice of a global
// scratch buffer that stays L2 resident (ViT-H/14: 257 tokens, 384-pixel models: 577; the reference handles any N).
// Exactly three iterations, no epsilon: that is the reference's contract (column sums are generally not 1 on exit).
// The same forward kernel also materialises the probabilities ([B,H,N,N] fp32, softmax or Sinkhorn) for the
// introspection path (recorder.py:28-31 hooks Attention.attend): nrv_attn_probs.
// Forward stashes lse and the 7 normalisation vectors ([B,H,8,N] fp32); backward rebuilds the final
// matrix from them (x_k = y_k * s_k walks the chain backwards without recomputing any reduction),
// keeps the gradient matrix in a per-CTA global scratch (L2 resident) and applies
//   y = x / s, s = sum_axis x   =>   dx = (dy - sum_axis(dy o y)) / s
// for the 7 steps in reverse, then the softmax backward.  Deterministic, no atomics.
#include "common.cuh"
#include "nrvit_internal.h"

namespace nrv {

constexpr int SK_WARPS = 16;
constexpr int SK_THREADS = SK_WARPS * 32;
constexpr int SK_STEPS = 7;   // row, col, row, col, row, col, row

__device__ __forceinline__ bool sk_is_row(int k) { return (k & 1) == 0; }   // k = 0..6

template <typename T>
__device__ __forceinline__ void sk_load_matrix(float* dst, const T* src, int N, int dh, int ldd, long long stride) {
  for (int i = threadIdx.x; i < N * dh; i += blockDim.x) {
    const int n = i / dh, d = i - n * dh;
    dst[n * ldd + d] = to_f32(src[(long long)n * stride + d]);
  }
}

// P[i][j] = exp(scale * q_i . k_j - lse_i) with K in shared memory; optionally computes lse
template <typename T>
__device__ __forceinline__ void sk_softmax_rows(float* P, int ldp, const float* Ks, int ldd, const T* qbase,
                                                long long tok_stride, int N, int dh, float scale, float* lse_io,
                                                bool compute_lse, float* qrow /* [warps][dh] */) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* q = qrow + warp * dh;
  for (int i = warp; i < N; i += SK_WARPS) {
    for (int d = lane; d < dh; d += 32) q[d] = to_f32(qbase[(long long)i * tok_stride + d]);
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) {
      float s = 0.f;
      for (int d = 0; d < dh; ++d) s = fmaf(q[d], Ks[j * ldd + d], s);
      s *= scale;
      P[i * ldp + j] = s;
      mx = fmaxf(mx, s);
    }
    float l;
    if (compute_lse) {
      mx = warp_max(mx);
      float sum = 0.f;
      for (int j = lane; j < N; j += 32) sum += expf(P[i * ldp + j] - mx);
      sum = warp_sum(sum);
      l = mx + logf(sum);
      if (lane == 0) lse_io[i] = l;
    } else {
      l = lse_io[i];
    }
    for (int j = lane; j < N; j += 32) P[i * ldp + j] = expf(P[i * ldp + j] - l);
    __syncwarp();
  }
}

template <typename T>
__global__ void __launch_bounds__(SK_THREADS, 1) sinkhorn_fwd_kernel(const T* __restrict__ qkv, T* __restrict__ out,
                                                                     float* __restrict__ stats, float* __restrict__ probs,
                                                                     float* __restrict__ scratch, int items, int N, int H,
                                                                     int dh, float scale, int n_steps) {
  extern __shared__ float sm[];
  const int ldd = dh + 1, ldp = N | 1;   // odd pitch: column walks are conflict free
  // P: shared memory, or (scratch != nullptr) this CTA's slice of the global scratch
  float* P = scratch ? scratch + (long long)blockIdx.x * N * ldp : sm;
  float* M = scratch ? sm : sm + N * ldp;   // K, later V   [N][dh+1]
  float* qrow = M + N * ldd;             // [warps][dh]
  float* vec = qrow + SK_WARPS * dh;     // [N] current sums
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long tok = 3ll * H * dh;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = item / H, h = item % H;
    const T* base = qkv + (long long)b * N * tok + (long long)h * dh;
    float* st = stats + (long long)item * (n_steps > 0 ? 8 : 1) * N;   // [8][N]: lse, then the 7 sum vectors ([1][N]: lse only)
    __syncthreads();   // previous item done with shared memory
    sk_load_matrix(M, base + (long long)H * dh, N, dh, ldd, tok);   // K
    __syncthreads();
    sk_softmax_rows(P, ldp, M, ldd, base, tok, N, dh, scale, st, true, qrow);
    __syncthreads();
    if (out != nullptr) sk_load_matrix(M, base + 2ll * H * dh, N, dh, ldd, tok);         // V (K is dead)
    for (int k = 0; k < n_steps; ++k) {
      if (sk_is_row(k)) {
        for (int i = warp; i < N; i += SK_WARPS) {
          float s = 0.f;
          for (int j = lane; j < N; j += 32) s += P[i * ldp + j];
          s = warp_sum(s);
          const float inv = 1.f / s;
          for (int j = lane; j < N; j += 32) P[i * ldp + j] *= inv;
          if (lane == 0) st[(1 + k) * N + i] = s;
        }
      } else {
        for (int j = threadIdx.x; j < N; j += SK_THREADS) {
          float s = 0.f;
          for (int i = 0; i < N; ++i) s += P[i * ldp + j];
          vec[j] = 1.f / s;
          st[(1 + k) * N + j] = s;
        }
        __syncthreads();
        for (int i = warp; i < N; i += SK_WARPS)
          for (int j = lane; j < N; j += 32) P[i * ldp + j] *= vec[j];
      }
      __syncthreads();
    }
    if (probs != nullptr) {   // introspection: the matrix `attend` returns
      float* po = probs + (long long)item * N * N;
      for (int i = warp; i < N; i += SK_WARPS)
        for (int j = lane; j < N; j += 32) po[(long long)i * N + j] = P[i * ldp + j];
    }
    if (out == nullptr) continue;
    // out = P V
    for (int i = warp; i < N; i += SK_WARPS) {
      for (int d = lane; d < dh; d += 32) {
        float acc = 0.f;
        for (int j = 0; j < N; ++j) acc = fmaf(P[i * ldp + j], M[j * ldd + d], acc);
        out[((long long)b * N + i) * H * dh + (long long)h * dh + d] = from_f32<T>(acc);
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(SK_THREADS, 1) sinkhorn_bwd_kernel(
    const T* __restrict__ qkv, const T* __restrict__ dout, const float* __restrict__ stats, T* __restrict__ dqkv,
    float* __restrict__ scratch, int items, int N, int H, int dh, float scale, int p_global) {
  extern __shared__ float sm[];
  const int ldd = dh + 1, ldp = N | 1;
  // gradient matrix of this CTA (row pitch N) always lives in the global scratch; P too when it does not fit on chip
  float* G = scratch + (long long)blockIdx.x * ((long long)N * N + (p_global ? (long long)N * ldp : 0));
  float* P = p_global ? G + (long long)N * N : sm;
  float* M = p_global ? sm : sm + N * ldp;   // K -> V -> K -> Q
  float* qrow = M + N * ldd;             // [warps][dh]
  float* vec = qrow + SK_WARPS * dh;     // [N]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long tok = 3ll * H * dh, os = (long long)H * dh;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = item / H, h = item % H;
    const T* base = qkv + (long long)b * N * tok + (long long)h * dh;
    const T* dob = dout + (long long)b * N * os + (long long)h * dh;
    T* dbase = dqkv + (long long)b * N * tok + (long long)h * dh;
    const float* st = stats + ((long long)b * H + h) * 8 * N;
    __syncthreads();   // previous item done with smem
    // ---- rebuild P0 = softmax(S) and the final matrix (divide by the stashed sums)
    sk_load_matrix(M, base + (long long)H * dh, N, dh, ldd, tok);   // K
    for (int i = threadIdx.x; i < N; i += SK_THREADS) vec[i] = st[i];   // lse
    __syncthreads();
    sk_softmax_rows(P, ldp, M, ldd, base, tok, N, dh, scale, vec, false, qrow);
    __syncthreads();
    for (int k = 0; k < SK_STEPS; ++k) {
      const float* s = st + (1 + k) * N;
      const bool row = sk_is_row(k);
      for (int i = warp; i < N; i += SK_WARPS) {
        const float si = row ? 1.f / s[i] : 0.f;
        for (int j = lane; j < N; j += 32) P[i * ldp + j] *= row ? si : 1.f / s[j];
      }
      __syncthreads();
    }
    // ---- dV_j = sum_i P_ij dO_i ;  G = dP = dO V^T
    sk_load_matrix(M, base + 2ll * H * dh, N, dh, ldd, tok);         // V
    __syncthreads();
    for (int j = warp; j < N; j += SK_WARPS) {
      for (int d = lane; d < dh; d += 32) {
        float acc = 0.f;
        for (int i = 0; i < N; ++i) acc = fmaf(P[i * ldp + j], to_f32(dob[(long long)i * os + d]), acc);
        dbase[(long long)j * tok + 2ll * H * dh + d] = from_f32<T>(acc);
      }
    }
    {
      float* q = qrow + warp * dh;
      for (int i = warp; i < N; i += SK_WARPS) {
        for (int d = lane; d < dh; d += 32) q[d] = to_f32(dob[(long long)i * os + d]);
        __syncwarp();
        for (int j = lane; j < N; j += 32) {
          float s = 0.f;
          for (int d = 0; d < dh; ++d) s = fmaf(q[d], M[j * ldd + d], s);
          G[(long long)i * N + j] = s;
        }
        __syncwarp();
      }
    }
    __syncthreads();
    // ---- the 7 normalisation steps in reverse: dx = (dy - sum_axis(dy o y)) / s ; x = y * s
    for (int k = SK_STEPS - 1; k >= 0; --k) {
      const float* s = st + (1 + k) * N;
      if (sk_is_row(k)) {
        for (int i = warp; i < N; i += SK_WARPS) {
          float t = 0.f;
          for (int j = lane; j < N; j += 32) t = fmaf(G[(long long)i * N + j], P[i * ldp + j], t);
          t = warp_sum(t);
          const float si = s[i], inv = 1.f / si;
          for (int j = lane; j < N; j += 32) {
            G[(long long)i * N + j] = (G[(long long)i * N + j] - t) * inv;
            P[i * ldp + j] *= si;
          }
        }
      } else {
        for (int j = threadIdx.x; j < N; j += SK_THREADS) {
          float t = 0.f;
          for (int i = 0; i < N; ++i) t = fmaf(G[(long long)i * N + j], P[i * ldp + j], t);
          vec[j] = t;
        }
        __syncthreads();
        for (int i = warp; i < N; i += SK_WARPS)
          for (int j = lane; j < N; j += 32) {
            const float sj = s[j];
            G[(long long)i * N + j] = (G[(long long)i * N + j] - vec[j]) / sj;
            P[i * ldp + j] *= sj;
          }
      }
      __syncthreads();
    }
    // ---- softmax backward: dS = P0 o (dP0 - rowsum(dP0 o P0)) * scale   (P now holds P0)
    for (int i = warp; i < N; i += SK_WARPS) {
      float t = 0.f;
      for (int j = lane; j < N; j += 32) t = fmaf(G[(long long)i * N + j], P[i * ldp + j], t);
      t = warp_sum(t);
      for (int j = lane; j < N; j += 32) P[i * ldp + j] = P[i * ldp + j] * (G[(long long)i * N + j] - t) * scale;
    }
    // ---- dQ_i = sum_j dS_ij K_j ; dK_j = sum_i dS_ij Q_i     (P now holds dS)
    sk_load_matrix(M, base + (long long)H * dh, N, dh, ldd, tok);   // K
    __syncthreads();
    for (int i = warp; i < N; i += SK_WARPS)
      for (int d = lane; d < dh; d += 32) {
        float acc = 0.f;
        for (int j = 0; j < N; ++j) acc = fmaf(P[i * ldp + j], M[j * ldd + d], acc);
        dbase[(long long)i * tok + d] = from_f32<T>(acc);
      }
    __syncthreads();
    sk_load_matrix(M, base, N, dh, ldd, tok);                        // Q
    __syncthreads();
    for (int j = warp; j < N; j += SK_WARPS)
      for (int d = lane; d < dh; d += 32) {
        float acc = 0.f;
        for (int i = 0; i < N; ++i) acc = fmaf(P[i * ldp + j], M[i * ldd + d], acc);
        dbase[(long long)j * tok + (long long)H * dh + d] = from_f32<T>(acc);
      }
  }
}

static size_t sk_smem_bytes(int N, int dh, bool p_global) {
  return ((p_global ? 0 : (size_t)N * (N | 1)) + (size_t)N * (dh + 1) + (size_t)SK_WARPS * dh + (size_t)N) * sizeof(float);
}
static const size_t SK_SMEM_MAX = (size_t)227 * 1024;

// on-chip probability matrix (fast path) vs L2-resident scratch; beyond that the K / V tile itself does not fit
static bool sk_fits_smem(int N, int dh) { return sk_smem_bytes(N, dh, false) <= SK_SMEM_MAX; }
bool sinkhorn_supported(int N, int dh) { return sk_smem_bytes(N, dh, true) <= SK_SMEM_MAX; }

static int sk_grid(int items) {
  const int sms = num_sms() > 0 ? num_sms() : 148;
  return items < sms ? items : sms;
}

size_t sinkhorn_fwd_scratch_bytes(int B, int N, int H, int dh) {
  if (sk_fits_smem(N, dh)) return 0;
  return (size_t)sk_grid(B * H) * N * (N | 1) * sizeof(float) + 256;
}

size_t sinkhorn_bwd_scratch_bytes(int B, int N, int H, int dh) {
  const size_t per_cta = (size_t)N * N + (sk_fits_smem(N, dh) ? 0 : (size_t)N * (N | 1));
  return (size_t)sk_grid(B * H) * per_cta * sizeof(float) + 256;
}

// out == nullptr: probabilities only (probs != nullptr).  n_steps = 7 (Sinkhorn) or 0 (plain softmax).
static int sk_launch_fwd(const void* qkv, void* out, float* stats, float* probs, void* scratch, size_t scratch_bytes,
                         int B, int N, int H, int dh, float scale, int dtype, int n_steps, cudaStream_t st) {
  if (!sinkhorn_supported(N, dh)) {
    set_error("Sinkhorn attention: N=%d dh=%d needs %zu bytes of shared memory for one K / V tile (max 227 KB)", N, dh,
              sk_smem_bytes(N, dh, true));
    return NRV_ENOTIMPL;
  }
  NRV_REQUIRE(stats != nullptr, "Sinkhorn attention needs the [B,H,8,N] fp32 statistics buffer");
  const bool p_global = !sk_fits_smem(N, dh);
  NRV_REQUIRE(!p_global || (scratch != nullptr && scratch_bytes >= sinkhorn_fwd_scratch_bytes(B, N, H, dh)),
              "Sinkhorn attention with N=%d tokens needs a scratch buffer of nrv_attn_fwd_workspace() bytes", N);
  const size_t smem = sk_smem_bytes(N, dh, p_global);
  float* sc = p_global ? reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(scratch) + 255) & ~uintptr_t(255)) : nullptr;
  const int items = B * H, grid = sk_grid(items);
  if (dtype == NRV_BF16) {
    NRV_CUDA(cudaFuncSetAttribute(sinkhorn_fwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SK_SMEM_MAX));
    sinkhorn_fwd_kernel<bf16><<<grid, SK_THREADS, smem, st>>>((const bf16*)qkv, (bf16*)out, stats, probs, sc, items, N, H, dh, scale, n_steps);
  } else {
    NRV_CUDA(cudaFuncSetAttribute(sinkhorn_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SK_SMEM_MAX));
    sinkhorn_fwd_kernel<float><<<grid, SK_THREADS, smem, st>>>((const float*)qkv, (float*)out, stats, probs, sc, items, N, H, dh, scale, n_steps);
  }
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

int sinkhorn_fwd(const void* qkv, void* out, float* stats, void* scratch, size_t scratch_bytes, int B, int N, int H, int dh,
                 float scale, int dtype, cudaStream_t st) {
  return sk_launch_fwd(qkv, out, stats, nullptr, scratch, scratch_bytes, B, N, H, dh, scale, dtype, SK_STEPS, st);
}

// probabilities of every head as the reference's `attend` module returns them (softmax, or softmax + Sinkhorn)
int attn_probs(const void* qkv, float* probs, float* stats, void* scratch, size_t scratch_bytes, int B, int N, int H, int dh,
               float scale, int sinkhorn, int dtype, cudaStream_t st) {
  NRV_REQUIRE(probs != nullptr, "nrv_attn_probs: null output");
  return sk_launch_fwd(qkv, nullptr, stats, probs, scratch, scratch_bytes, B, N, H, dh, scale, dtype, sinkhorn ? SK_STEPS : 0, st);
}

int sinkhorn_bwd(const void* qkv, const void* dout, const float* stats, void* dqkv, float* scratch, int B, int N,
                 int H, int dh, float scale, int dtype, cudaStream_t st) {
  if (!sinkhorn_supported(N, dh)) {
    set_error("Sinkhorn attention: N=%d dh=%d needs %zu bytes of shared memory for one K / V tile (max 227 KB)", N, dh,
              sk_smem_bytes(N, dh, true));
    return NRV_ENOTIMPL;
  }
  const int p_global = sk_fits_smem(N, dh) ? 0 : 1;
  const size_t smem = sk_smem_bytes(N, dh, p_global != 0);
  const int items = B * H;
  const int grid = sk_grid(items);
  scratch = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(scratch) + 255) & ~uintptr_t(255));
  if (dtype == NRV_BF16) {
    NRV_CUDA(cudaFuncSetAttribute(sinkhorn_bwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SK_SMEM_MAX));
    sinkhorn_bwd_kernel<bf16><<<grid, SK_THREADS, smem, st>>>((const bf16*)qkv, (const bf16*)dout, stats, (bf16*)dqkv,
                                                              scratch, items, N, H, dh, scale, p_global);
  } else {
    NRV_CUDA(cudaFuncSetAttribute(sinkhorn_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SK_SMEM_MAX));
    sinkhorn_bwd_kernel<float><<<grid, SK_THREADS, smem, st>>>((const float*)qkv, (const float*)dout, stats,
                                                               (float*)dqkv, scratch, items, N, H, dh, scale, p_global);
  }
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

}  // namespace nrv
