// fp32 CUDA-core attention (forward + backward), one CTA per (batch, head).
//
// Role: (1) the attention of the NRV_F32 check mode (fp32 activations, exact fp32 FMA arithmetic),
//       (2) an independent on-device cross-check of the tcgen05 attention kernel, and
//       (3) the fallback for shapes the tcgen05 kernel does not cover.
// It is NOT the production bf16 path (see attention_tc.cu).
//
// Reference semantics: simple_vit.py:70-75 — dots = q k^T * scale ; attn = softmax(dots, -1) ;
// out = attn v ; 'b h n d -> b n (h d)'.  qkv is the packed projection output
// [B, N, 3, H, dh] (q|k|v, then head-major: simple_vit.py:67-68 ; utils.py:115,489-502).
// Backward (autograd of the above): with P = exp(S - lse),  dV = P^T dO ; dP = dO V^T ;
// delta = rowsum(dO o O) ; dS = P o (dP - delta) ; dQ = scale dS K ; dK = scale dS^T Q.
// Deterministic: no atomics, fixed summation order.
#include "common.cuh"
#include "nrvit_internal.h"

namespace nrv {

constexpr int SIMT_WARPS = 8;

template <typename T>
__device__ __forceinline__ void load_head_matrix(float* dst, const T* src, int N, int dh, int ldd,
                                                 long long row_stride) {
  // dst[n][d] (row stride ldd) <- src[n*row_stride + d]
  for (int i = threadIdx.x; i < N * dh; i += blockDim.x) {
    const int n = i / dh, d = i - n * dh;
    dst[n * ldd + d] = to_f32(src[(long long)n * row_stride + d]);
  }
}

template <typename T>
__global__ void __launch_bounds__(SIMT_WARPS * 32) attn_fwd_simt_kernel(
    const T* __restrict__ qkv, T* __restrict__ out, float* __restrict__ lse, int N, int H, int dh,
    float scale) {
  extern __shared__ float sm[];
  const int ldd = dh + 1;
  float* Ks = sm;
  float* Vs = Ks + N * ldd;
  float* qs = Vs + N * ldd;             // [warps][dh]
  float* ps = qs + SIMT_WARPS * dh;     // [warps][N]
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long tok_stride = 3ll * H * dh;
  const T* base = qkv + (long long)b * N * tok_stride + (long long)h * dh;
  load_head_matrix(Ks, base + (long long)H * dh, N, dh, ldd, tok_stride);
  load_head_matrix(Vs, base + 2ll * H * dh, N, dh, ldd, tok_stride);
  __syncthreads();
  float* q = qs + warp * dh;
  float* p = ps + warp * N;
  for (int i = warp; i < N; i += SIMT_WARPS) {
    for (int d = lane; d < dh; d += 32) q[d] = to_f32(base[(long long)i * tok_stride + d]);
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) {
      float s = 0.f;
      for (int d = 0; d < dh; ++d) s = fmaf(q[d], Ks[j * ldd + d], s);
      s *= scale;
      p[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < N; j += 32) {
      const float e = expf(p[j] - mx);
      p[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    if (lane == 0 && lse) lse[((long long)b * H + h) * N + i] = mx + logf(sum);
    __syncwarp();
    for (int d = lane; d < dh; d += 32) {
      float acc = 0.f;
      for (int j = 0; j < N; ++j) acc = fmaf(p[j], Vs[j * ldd + d], acc);
      out[((long long)b * N + i) * H * dh + (long long)h * dh + d] = from_f32<T>(acc * inv);
    }
    __syncwarp();
  }
}

template <typename T>
__global__ void __launch_bounds__(SIMT_WARPS * 32) attn_bwd_simt_kernel(
    const T* __restrict__ qkv, const T* __restrict__ out, const T* __restrict__ dout,
    const float* __restrict__ lse, T* __restrict__ dqkv, int N, int H, int dh, float scale) {
  extern __shared__ float sm[];
  const int ldd = dh + 1;
  float* Qs = sm;
  float* Ks = Qs + N * ldd;
  float* Vs = Ks + N * ldd;
  float* Ds = Vs + N * ldd;             // dO
  float* ls = Ds + N * ldd;             // lse   [N]
  float* dl = ls + N;                   // delta [N]
  float* pa = dl + N;                   // [warps][N]
  float* pb = pa + SIMT_WARPS * N;      // [warps][N]
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long tok_stride = 3ll * H * dh;
  const long long o_stride = (long long)H * dh;
  const T* base = qkv + (long long)b * N * tok_stride + (long long)h * dh;
  const T* obase = out + (long long)b * N * o_stride + (long long)h * dh;
  const T* dobase = dout + (long long)b * N * o_stride + (long long)h * dh;
  T* dbase = dqkv + (long long)b * N * tok_stride + (long long)h * dh;
  load_head_matrix(Qs, base, N, dh, ldd, tok_stride);
  load_head_matrix(Ks, base + (long long)H * dh, N, dh, ldd, tok_stride);
  load_head_matrix(Vs, base + 2ll * H * dh, N, dh, ldd, tok_stride);
  load_head_matrix(Ds, dobase, N, dh, ldd, o_stride);
  for (int i = threadIdx.x; i < N; i += blockDim.x) ls[i] = lse[((long long)b * H + h) * N + i];
  __syncthreads();
  // delta_i = sum_d dO[i][d] * O[i][d]
  for (int i = warp; i < N; i += SIMT_WARPS) {
    float s = 0.f;
    for (int d = lane; d < dh; d += 32) s = fmaf(Ds[i * ldd + d], to_f32(obase[(long long)i * o_stride + d]), s);
    s = warp_sum(s);
    if (lane == 0) dl[i] = s;
  }
  __syncthreads();
  float* wa = pa + warp * N;
  float* wb = pb + warp * N;
  // phase A: dQ_i = scale * sum_j dS_ij K_j
  for (int i = warp; i < N; i += SIMT_WARPS) {
    const float li = ls[i], di = dl[i];
    for (int j = lane; j < N; j += 32) {
      float s = 0.f, dp = 0.f;
      for (int d = 0; d < dh; ++d) {
        s = fmaf(Qs[i * ldd + d], Ks[j * ldd + d], s);
        dp = fmaf(Ds[i * ldd + d], Vs[j * ldd + d], dp);
      }
      const float p = expf(s * scale - li);
      wa[j] = p * (dp - di) * scale;
    }
    __syncwarp();
    for (int d = lane; d < dh; d += 32) {
      float acc = 0.f;
      for (int j = 0; j < N; ++j) acc = fmaf(wa[j], Ks[j * ldd + d], acc);
      dbase[(long long)i * tok_stride + d] = from_f32<T>(acc);
    }
    __syncwarp();
  }
  // phase B: dK_j = scale * sum_i dS_ij Q_i ; dV_j = sum_i P_ij dO_i
  for (int j = warp; j < N; j += SIMT_WARPS) {
    for (int i = lane; i < N; i += 32) {
      float s = 0.f, dp = 0.f;
      for (int d = 0; d < dh; ++d) {
        s = fmaf(Qs[i * ldd + d], Ks[j * ldd + d], s);
        dp = fmaf(Ds[i * ldd + d], Vs[j * ldd + d], dp);
      }
      const float p = expf(s * scale - ls[i]);
      wa[i] = p;
      wb[i] = p * (dp - dl[i]) * scale;
    }
    __syncwarp();
    for (int d = lane; d < dh; d += 32) {
      float ak = 0.f, av = 0.f;
      for (int i = 0; i < N; ++i) {
        ak = fmaf(wb[i], Qs[i * ldd + d], ak);
        av = fmaf(wa[i], Ds[i * ldd + d], av);
      }
      dbase[(long long)j * tok_stride + (long long)H * dh + d] = from_f32<T>(ak);
      dbase[(long long)j * tok_stride + 2ll * H * dh + d] = from_f32<T>(av);
    }
    __syncwarp();
  }
}

static const int kMaxSmem = 227 * 1024;

int attn_fwd_simt(const void* qkv, void* out, float* lse, int B, int N, int H, int dh, float scale,
                  int dtype, cudaStream_t st) {
  const size_t smem = ((size_t)2 * N * (dh + 1) + (size_t)SIMT_WARPS * (dh + N)) * sizeof(float);
  if (smem > (size_t)kMaxSmem) {
    set_error("nrv_attn_fwd(SIMT): N=%d dh=%d needs %zu bytes of shared memory (> %d)", N, dh, smem, kMaxSmem);
    return NRV_ENOTIMPL;
  }
  if (dtype == NRV_BF16) {
    NRV_CUDA(cudaFuncSetAttribute(attn_fwd_simt_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_fwd_simt_kernel<bf16><<<B * H, SIMT_WARPS * 32, smem, st>>>((const bf16*)qkv, (bf16*)out, lse, N, H, dh, scale);
  } else {
    NRV_CUDA(cudaFuncSetAttribute(attn_fwd_simt_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_fwd_simt_kernel<float><<<B * H, SIMT_WARPS * 32, smem, st>>>((const float*)qkv, (float*)out, lse, N, H, dh, scale);
  }
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

int attn_bwd_simt(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                  int B, int N, int H, int dh, float scale, int dtype, cudaStream_t st) {
  const size_t smem = ((size_t)4 * N * (dh + 1) + 2 * (size_t)N + 2 * (size_t)SIMT_WARPS * N) * sizeof(float);
  if (smem > (size_t)kMaxSmem) {
    set_error("nrv_attn_bwd(SIMT): N=%d dh=%d needs %zu bytes of shared memory (> %d)", N, dh, smem, kMaxSmem);
    return NRV_ENOTIMPL;
  }
  if (dtype == NRV_BF16) {
    NRV_CUDA(cudaFuncSetAttribute(attn_bwd_simt_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_bwd_simt_kernel<bf16><<<B * H, SIMT_WARPS * 32, smem, st>>>((const bf16*)qkv, (const bf16*)out, (const bf16*)dout, lse, (bf16*)dqkv, N, H, dh, scale);
  } else {
    NRV_CUDA(cudaFuncSetAttribute(attn_bwd_simt_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_bwd_simt_kernel<float><<<B * H, SIMT_WARPS * 32, smem, st>>>((const float*)qkv, (const float*)out, (const float*)dout, lse, (float*)dqkv, N, H, dh, scale);
  }
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

}  // namespace nrv
