// fp32 CUDA-core attention (forward + backward), one CTA per (batch, head).
//
// Role: (1) the attention of the NRV_F32 check mode (fp32 activations, exact fp32 FMA arithmetic),
//       (2) an independent on-device cross-check of the tcgen05 attention kernel, and
//       (3) the fallback for shapes the tcgen05 kernel does not cover.
// It is NOT the production bf16 path (see attention_fwd2.cu / attention_bwd2.cu).
//
// Reference semantics: simple_vit.py:70-75 — dots = q k^T * scale ; attn = softmax(dots, -1) ;
// out = attn v ; 'b h n d -> b n (h d)'.  qkv is the packed projection output
// [B, N, 3, H, dh] (q|k|v, then head-major: simple_vit.py:67-68 ; utils.py:115,489-502).
// Backward (autograd of the above): with P = exp(S - lse),  dV = P^T dO ; dP = dO V^T ;
// delta = rowsum(dO o O) ; dS = P o (dP - delta) ; dQ = scale dS K ; dK = scale dS^T Q.
// Deterministic: no atomics, fixed summation order.
// Dropout on the attention probabilities (nn.MultiheadAttention dropout, vit.py:105-110 ; README ViT Attention.dropout):
// out = (P o M) V with M_ij = keep_ij / (1 - p), keep_ij the counter-based decision of nrv_dropout for element
// ((b*H + h)*N + i)*N + j at site NRV_DROP_ATTN_PROB.  Backward: dV = (P o M)^T dO ; dP = (dO V^T) o M ;
// delta = rowsum(dO o O) still equals rowsum(P o dP).  Only these kernels implement it (the tcgen05 kernels do not).
#include "common.cuh"
#include "nrvit_internal.h"

namespace nrv {

constexpr int SIMT_WARPS = 8;

// same Philox4x32-10 stream as dropout_kernel (elementwise.cu): call e / 4, component e % 4 (attn_keep4, common.cuh)
__device__ __forceinline__ float attn_drop_factor(const AttnDrop& dr, uint32_t thresh, float keep_scale, unsigned long long e) {
  return ((attn_keep4(dr, thresh, e >> 2) >> ((uint32_t)e & 3u)) & 1u) ? keep_scale : 0.f;
}

template <typename T, typename ST>
__device__ __forceinline__ void load_head_matrix(ST* dst, const T* src, int N, int dh, int ldd,
                                                 long long row_stride) {
  // dst[n][d] (row stride ldd) <- src[n*row_stride + d]
  for (int i = threadIdx.x; i < N * dh; i += blockDim.x) {
    const int n = i / dh, d = i - n * dh;
    dst[n * ldd + d] = from_f32<ST>(to_f32(src[(long long)n * row_stride + d]));
  }
}

// Forward: a warp works on RQ = 4 query rows at a time, so every K / V element read from shared memory feeds four
// FMAs (q and p of the four rows sit interleaved: one broadcast LDS.128 per step) instead of one -- this kernel is
// also the bf16 fallback for shapes the tcgen05 kernel does not cover (ViT-H/14: dh = 80, N = 257).
constexpr int SIMT_RQ = 4;

template <typename T>
__global__ void __launch_bounds__(SIMT_WARPS * 32) attn_fwd_simt_kernel(
    const T* __restrict__ qkv, T* __restrict__ out, float* __restrict__ lse, int N, int H, int dh,
    float scale, AttnDrop dr) {
  extern __shared__ float sm[];
  const uint32_t dthresh = (uint32_t)(dr.p * 16777216.0f);
  const float dscale = 1.f / (1.f - dr.p);
  const int ldd = dh + 1;
  float* Ks = sm;
  float* Vs = Ks + N * ldd;
  float4* qs = reinterpret_cast<float4*>(Vs + N * ldd + ((4 - ((2 * N * ldd) & 3)) & 3));   // [warps][dh] x 4 rows, 16-byte aligned
  float4* ps = qs + SIMT_WARPS * dh;                                                       // [warps][N]  x 4 rows
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long tok_stride = 3ll * H * dh;
  const T* base = qkv + (long long)b * N * tok_stride + (long long)h * dh;
  load_head_matrix(Ks, base + (long long)H * dh, N, dh, ldd, tok_stride);
  load_head_matrix(Vs, base + 2ll * H * dh, N, dh, ldd, tok_stride);
  __syncthreads();
  float4* q = qs + warp * dh;
  float4* p = ps + warp * N;
  for (int i0 = warp * SIMT_RQ; i0 < N; i0 += SIMT_WARPS * SIMT_RQ) {
    for (int d = lane; d < dh; d += 32) {
      float v[SIMT_RQ];
#pragma unroll
      for (int r = 0; r < SIMT_RQ; ++r) v[r] = i0 + r < N ? to_f32(base[(long long)(i0 + r) * tok_stride + d]) : 0.f;
      q[d] = make_float4(v[0], v[1], v[2], v[3]);
    }
    __syncwarp();
    float mx[SIMT_RQ] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int j = lane; j < N; j += 32) {
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      const float* kr = Ks + j * ldd;
      for (int d = 0; d < dh; ++d) {
        const float k = kr[d];
        const float4 qq = q[d];
        s0 = fmaf(qq.x, k, s0); s1 = fmaf(qq.y, k, s1); s2 = fmaf(qq.z, k, s2); s3 = fmaf(qq.w, k, s3);
      }
      s0 *= scale; s1 *= scale; s2 *= scale; s3 *= scale;
      p[j] = make_float4(s0, s1, s2, s3);
      mx[0] = fmaxf(mx[0], s0); mx[1] = fmaxf(mx[1], s1); mx[2] = fmaxf(mx[2], s2); mx[3] = fmaxf(mx[3], s3);
    }
#pragma unroll
    for (int r = 0; r < SIMT_RQ; ++r) mx[r] = warp_max(mx[r]);
    float sum[SIMT_RQ] = {0.f, 0.f, 0.f, 0.f};
    for (int j = lane; j < N; j += 32) {
      float4 e = p[j];
      e.x = expf(e.x - mx[0]); e.y = expf(e.y - mx[1]); e.z = expf(e.z - mx[2]); e.w = expf(e.w - mx[3]);
      sum[0] += e.x; sum[1] += e.y; sum[2] += e.z; sum[3] += e.w;
      if (dr.p > 0.f) {   // the row sum is over the undropped probabilities; P V sees the masked ones
        const unsigned long long e0 = ((unsigned long long)blockIdx.x * N + i0) * N + j;
        e.x *= attn_drop_factor(dr, dthresh, dscale, e0);
        e.y *= attn_drop_factor(dr, dthresh, dscale, e0 + N);
        e.z *= attn_drop_factor(dr, dthresh, dscale, e0 + 2ull * N);
        e.w *= attn_drop_factor(dr, dthresh, dscale, e0 + 3ull * N);
      }
      p[j] = e;
    }
    float inv[SIMT_RQ];
#pragma unroll
    for (int r = 0; r < SIMT_RQ; ++r) {
      sum[r] = warp_sum(sum[r]);
      inv[r] = 1.f / sum[r];
      if (lane == 0 && lse && i0 + r < N) lse[((long long)b * H + h) * N + i0 + r] = mx[r] + logf(sum[r]);
    }
    __syncwarp();
    for (int d = lane; d < dh; d += 32) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int j = 0; j < N; ++j) {
        const float v = Vs[j * ldd + d];
        const float4 pp = p[j];
        a0 = fmaf(pp.x, v, a0); a1 = fmaf(pp.y, v, a1); a2 = fmaf(pp.z, v, a2); a3 = fmaf(pp.w, v, a3);
      }
      const float acc[SIMT_RQ] = {a0 * inv[0], a1 * inv[1], a2 * inv[2], a3 * inv[3]};
#pragma unroll
      for (int r = 0; r < SIMT_RQ; ++r)
        if (i0 + r < N) out[((long long)b * N + i0 + r) * H * dh + (long long)h * dh + d] = from_f32<T>(acc[r]);
    }
    __syncwarp();
  }
}

// ST = element type of the four head matrices in shared memory: float, or bf16 (exact for bf16 inputs) when the fp32
// copies do not fit -- e.g. ViT-H/14 training: 257 tokens, dh = 80.
template <typename T, typename ST>
__global__ void __launch_bounds__(SIMT_WARPS * 32) attn_bwd_simt_kernel(
    const T* __restrict__ qkv, const T* __restrict__ out, const T* __restrict__ dout,
    const float* __restrict__ lse, T* __restrict__ dqkv, int N, int H, int dh, float scale, AttnDrop dr) {
  extern __shared__ float sm[];
  const uint32_t dthresh = (uint32_t)(dr.p * 16777216.0f);
  const float dscale = 1.f / (1.f - dr.p);
  const int ldd = dh + 1;
  ST* Qs = reinterpret_cast<ST*>(sm);
  ST* Ks = Qs + N * ldd;
  ST* Vs = Ks + N * ldd;
  ST* Ds = Vs + N * ldd;                // dO
  float* ls = reinterpret_cast<float*>(Ds + N * ldd);   // lse [N]: 4 * N * ldd elements precede it, 8-byte aligned for any ST
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
    for (int d = lane; d < dh; d += 32) s = fmaf(to_f32(Ds[i * ldd + d]), to_f32(obase[(long long)i * o_stride + d]), s);
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
        s = fmaf(to_f32(Qs[i * ldd + d]), to_f32(Ks[j * ldd + d]), s);
        dp = fmaf(to_f32(Ds[i * ldd + d]), to_f32(Vs[j * ldd + d]), dp);
      }
      const float p = expf(s * scale - li);
      if (dr.p > 0.f) dp *= attn_drop_factor(dr, dthresh, dscale, ((unsigned long long)blockIdx.x * N + i) * N + j);
      wa[j] = p * (dp - di) * scale;
    }
    __syncwarp();
    for (int d = lane; d < dh; d += 32) {
      float acc = 0.f;
      for (int j = 0; j < N; ++j) acc = fmaf(wa[j], to_f32(Ks[j * ldd + d]), acc);
      dbase[(long long)i * tok_stride + d] = from_f32<T>(acc);
    }
    __syncwarp();
  }
  // phase B: dK_j = scale * sum_i dS_ij Q_i ; dV_j = sum_i P_ij dO_i
  for (int j = warp; j < N; j += SIMT_WARPS) {
    for (int i = lane; i < N; i += 32) {
      float s = 0.f, dp = 0.f;
      for (int d = 0; d < dh; ++d) {
        s = fmaf(to_f32(Qs[i * ldd + d]), to_f32(Ks[j * ldd + d]), s);
        dp = fmaf(to_f32(Ds[i * ldd + d]), to_f32(Vs[j * ldd + d]), dp);
      }
      const float p = expf(s * scale - ls[i]);
      float m = 1.f;
      if (dr.p > 0.f) m = attn_drop_factor(dr, dthresh, dscale, ((unsigned long long)blockIdx.x * N + i) * N + j);
      wa[i] = p * m;
      wb[i] = p * (dp * m - dl[i]) * scale;
    }
    __syncwarp();
    for (int d = lane; d < dh; d += 32) {
      float ak = 0.f, av = 0.f;
      for (int i = 0; i < N; ++i) {
        ak = fmaf(wb[i], to_f32(Qs[i * ldd + d]), ak);
        av = fmaf(wa[i], to_f32(Ds[i * ldd + d]), av);
      }
      dbase[(long long)j * tok_stride + (long long)H * dh + d] = from_f32<T>(ak);
      dbase[(long long)j * tok_stride + 2ll * H * dh + d] = from_f32<T>(av);
    }
    __syncwarp();
  }
}

// Same arithmetic for token counts whose four head matrices do not fit in shared memory together (fp32 check mode at
// ViT-H/14's 257 tokens x 80: 4 x 83 KB): only TWO matrices are resident at a time -- K and V while the warps walk the query
// rows (dQ), then Q and dO while they walk the key rows (dK, dV) -- and the row a warp works on comes from global memory
// into a small per-warp buffer.  Reach: 2 N (dh + 1) floats + vectors <= 227 KB (about 350 tokens at dh = 80).
template <typename T>
__global__ void __launch_bounds__(SIMT_WARPS * 32) attn_bwd_simt_stream_kernel(
    const T* __restrict__ qkv, const T* __restrict__ out, const T* __restrict__ dout,
    const float* __restrict__ lse, T* __restrict__ dqkv, int N, int H, int dh, float scale, AttnDrop dr) {
  extern __shared__ float sm[];
  const uint32_t dthresh = (uint32_t)(dr.p * 16777216.0f);
  const float dscale = 1.f / (1.f - dr.p);
  const int ldd = dh + 1;
  float* M0 = sm;                        // K, then Q
  float* M1 = M0 + N * ldd;              // V, then dO
  float* rows = M1 + N * ldd;            // [warps][2][ldd]: the row pair of the warp's current token
  float* ls = rows + SIMT_WARPS * 2 * ldd;
  float* dl = ls + N;
  float* pa = dl + N;                    // [warps][N]
  float* pb = pa + SIMT_WARPS * N;       // [warps][N]
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long tok_stride = 3ll * H * dh;
  const long long o_stride = (long long)H * dh;
  const T* base = qkv + (long long)b * N * tok_stride + (long long)h * dh;
  const T* obase = out + (long long)b * N * o_stride + (long long)h * dh;
  const T* dobase = dout + (long long)b * N * o_stride + (long long)h * dh;
  T* dbase = dqkv + (long long)b * N * tok_stride + (long long)h * dh;
  float* r0 = rows + warp * 2 * ldd;
  float* r1 = r0 + ldd;
  float* wa = pa + warp * N;
  float* wb = pb + warp * N;
  load_head_matrix(M0, base + (long long)H * dh, N, dh, ldd, tok_stride);       // K
  load_head_matrix(M1, base + 2ll * H * dh, N, dh, ldd, tok_stride);            // V
  for (int i = threadIdx.x; i < N; i += blockDim.x) ls[i] = lse[((long long)b * H + h) * N + i];
  // delta_i = sum_d dO[i][d] * O[i][d]
  for (int i = warp; i < N; i += SIMT_WARPS) {
    float s = 0.f;
    for (int d = lane; d < dh; d += 32) s = fmaf(to_f32(dobase[(long long)i * o_stride + d]), to_f32(obase[(long long)i * o_stride + d]), s);
    s = warp_sum(s);
    if (lane == 0) dl[i] = s;
  }
  __syncthreads();
  // phase A: dQ_i = scale * sum_j dS_ij K_j     (r0 = Q_i, r1 = dO_i)
  for (int i = warp; i < N; i += SIMT_WARPS) {
    for (int d = lane; d < dh; d += 32) {
      r0[d] = to_f32(base[(long long)i * tok_stride + d]);
      r1[d] = to_f32(dobase[(long long)i * o_stride + d]);
    }
    __syncwarp();
    const float li = ls[i], di = dl[i];
    for (int j = lane; j < N; j += 32) {
      float s = 0.f, dp = 0.f;
      for (int d = 0; d < dh; ++d) {
        s = fmaf(r0[d], M0[j * ldd + d], s);
        dp = fmaf(r1[d], M1[j * ldd + d], dp);
      }
      const float p = expf(s * scale - li);
      if (dr.p > 0.f) dp *= attn_drop_factor(dr, dthresh, dscale, ((unsigned long long)blockIdx.x * N + i) * N + j);
      wa[j] = p * (dp - di) * scale;
    }
    __syncwarp();
    for (int d = lane; d < dh; d += 32) {
      float acc = 0.f;
      for (int j = 0; j < N; ++j) acc = fmaf(wa[j], M0[j * ldd + d], acc);
      dbase[(long long)i * tok_stride + d] = from_f32<T>(acc);
    }
    __syncwarp();
  }
  __syncthreads();
  load_head_matrix(M0, base, N, dh, ldd, tok_stride);                            // Q
  load_head_matrix(M1, dobase, N, dh, ldd, o_stride);                            // dO
  __syncthreads();
  // phase B: dK_j = scale * sum_i dS_ij Q_i ; dV_j = sum_i P_ij dO_i     (r0 = K_j, r1 = V_j)
  for (int j = warp; j < N; j += SIMT_WARPS) {
    for (int d = lane; d < dh; d += 32) {
      r0[d] = to_f32(base[(long long)j * tok_stride + (long long)H * dh + d]);
      r1[d] = to_f32(base[(long long)j * tok_stride + 2ll * H * dh + d]);
    }
    __syncwarp();
    for (int i = lane; i < N; i += 32) {
      float s = 0.f, dp = 0.f;
      for (int d = 0; d < dh; ++d) {
        s = fmaf(M0[i * ldd + d], r0[d], s);
        dp = fmaf(M1[i * ldd + d], r1[d], dp);
      }
      const float p = expf(s * scale - ls[i]);
      float m = 1.f;
      if (dr.p > 0.f) m = attn_drop_factor(dr, dthresh, dscale, ((unsigned long long)blockIdx.x * N + i) * N + j);
      wa[i] = p * m;
      wb[i] = p * (dp * m - dl[i]) * scale;
    }
    __syncwarp();
    for (int d = lane; d < dh; d += 32) {
      float ak = 0.f, av = 0.f;
      for (int i = 0; i < N; ++i) {
        ak = fmaf(wb[i], M0[i * ldd + d], ak);
        av = fmaf(wa[i], M1[i * ldd + d], av);
      }
      dbase[(long long)j * tok_stride + (long long)H * dh + d] = from_f32<T>(ak);
      dbase[(long long)j * tok_stride + 2ll * H * dh + d] = from_f32<T>(av);
    }
    __syncwarp();
  }
}

static const int kMaxSmem = 227 * 1024;

int attn_fwd_simt(const void* qkv, void* out, float* lse, int B, int N, int H, int dh, float scale,
                  int dtype, cudaStream_t st, float p_drop, unsigned long long seed, int layer) {
  const AttnDrop dr = attn_make_drop(p_drop, seed, layer);
  const size_t smem = ((size_t)2 * N * (dh + 1) + 4 + (size_t)SIMT_WARPS * SIMT_RQ * (dh + N)) * sizeof(float);
  if (smem > (size_t)kMaxSmem) {
    set_error("nrv_attn_fwd(SIMT): N=%d dh=%d needs %zu bytes of shared memory (> %d)", N, dh, smem, kMaxSmem);
    return NRV_ENOTIMPL;
  }
  if (dtype == NRV_BF16) {
    NRV_CUDA(cudaFuncSetAttribute(attn_fwd_simt_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_fwd_simt_kernel<bf16><<<B * H, SIMT_WARPS * 32, smem, st>>>((const bf16*)qkv, (bf16*)out, lse, N, H, dh, scale, dr);
  } else {
    NRV_CUDA(cudaFuncSetAttribute(attn_fwd_simt_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_fwd_simt_kernel<float><<<B * H, SIMT_WARPS * 32, smem, st>>>((const float*)qkv, (float*)out, lse, N, H, dh, scale, dr);
  }
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

int attn_bwd_simt(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                  int B, int N, int H, int dh, float scale, int dtype, cudaStream_t st, float p_drop,
                  unsigned long long seed, int layer) {
  const AttnDrop dr = attn_make_drop(p_drop, seed, layer);
  const size_t tail = (2 * (size_t)N + 2 * (size_t)SIMT_WARPS * N) * sizeof(float) + 8;
  const size_t smem32 = (size_t)4 * N * (dh + 1) * sizeof(float) + tail;
  const size_t smem16 = (size_t)4 * N * (dh + 1) * sizeof(bf16) + tail;
  const size_t smem_stream = ((size_t)2 * N * (dh + 1) + (size_t)SIMT_WARPS * 2 * (dh + 1)) * sizeof(float) + tail;
  if (smem32 <= (size_t)kMaxSmem) {
    if (dtype == NRV_BF16) {
      NRV_CUDA(cudaFuncSetAttribute(attn_bwd_simt_kernel<bf16, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem32));
      attn_bwd_simt_kernel<bf16, float><<<B * H, SIMT_WARPS * 32, smem32, st>>>((const bf16*)qkv, (const bf16*)out, (const bf16*)dout, lse, (bf16*)dqkv, N, H, dh, scale, dr);
    } else {
      NRV_CUDA(cudaFuncSetAttribute(attn_bwd_simt_kernel<float, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem32));
      attn_bwd_simt_kernel<float, float><<<B * H, SIMT_WARPS * 32, smem32, st>>>((const float*)qkv, (const float*)out, (const float*)dout, lse, (float*)dqkv, N, H, dh, scale, dr);
    }
  } else if (dtype == NRV_BF16 && smem16 <= (size_t)kMaxSmem) {
    // bf16 inputs: the head matrices are kept as bf16 in shared memory (exact), which doubles the reach of this kernel
    NRV_CUDA(cudaFuncSetAttribute(attn_bwd_simt_kernel<bf16, bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
    attn_bwd_simt_kernel<bf16, bf16><<<B * H, SIMT_WARPS * 32, smem16, st>>>((const bf16*)qkv, (const bf16*)out, (const bf16*)dout, lse, (bf16*)dqkv, N, H, dh, scale, dr);
  } else if (smem_stream <= (size_t)kMaxSmem) {
    // two head matrices resident at a time (the fp32 check mode beyond ~215 tokens: ViT-H/14)
    if (dtype == NRV_BF16) {
      NRV_CUDA(cudaFuncSetAttribute(attn_bwd_simt_stream_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_stream));
      attn_bwd_simt_stream_kernel<bf16><<<B * H, SIMT_WARPS * 32, smem_stream, st>>>((const bf16*)qkv, (const bf16*)out, (const bf16*)dout, lse, (bf16*)dqkv, N, H, dh, scale, dr);
    } else {
      NRV_CUDA(cudaFuncSetAttribute(attn_bwd_simt_stream_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_stream));
      attn_bwd_simt_stream_kernel<float><<<B * H, SIMT_WARPS * 32, smem_stream, st>>>((const float*)qkv, (const float*)out, (const float*)dout, lse, (float*)dqkv, N, H, dh, scale, dr);
    }
  } else {
    set_error("nrv_attn_bwd(SIMT): N=%d dh=%d needs %zu bytes of shared memory (> %d)", N, dh, smem_stream, kMaxSmem);
    return NRV_ENOTIMPL;
  }
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

}  // namespace nrv
