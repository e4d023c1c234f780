// LayerNorm folded into the projection GEMM that consumes it (the "fused LayerNorm + QKV / FC1 GEMM" of the forward
// pass; reference ops simple_vit.py:65-67 (Attention.norm + to_qkv), :38-39 (FeedForward net.0 + net.1); vit.py:123,128).
//
//   LayerNorm(x)_k = (x_k - mu) rstd gamma_k + beta_k
//   y_n = sum_k LayerNorm(x)_k W_nk + b_n = rstd sum_k (x_k - mu) gamma_k W_nk + c_n ,   c_n = sum_k beta_k W_nk + b_n
// and, because sum_k (x_k - mu) = 0, any constant may be subtracted from the row gamma o W[n,:]:
//   sum_k (x_k - mu) gamma_k W_nk = sum_k x_k Wc_nk      with  Wc_nk = gamma_k W_nk - mean_k(gamma o W[n,:])   (rows sum to zero)
// so the GEMM runs on the RAW residual stream with B = Wc and its epilogue is the ordinary bias epilogue with a per-row
// scale rstd_m (gemm.cu: ln_stats) -- no rank-one correction per element.  For the mean of x to cancel inside the
// accumulation the STORED (rounded) row must sum to zero; the fold kernel re-rounds one small element per row so that it
// does to ~2^-8 of the remainder (bf16: |row sum| ~1e-6 instead of ~1e-3), which keeps the result independent of
// |mu| / sigma.  The row statistics come for free from the epilogue of the GEMM that produced the row (stats_out:
// out-proj / FC2 + residual); only the embedding output needs the small reduction kernel below.
// No normalised copy of the stream is written or read in the forward pass; a backward pass, which needs it as the
// B operand of the weight-gradient GEMMs, gets it as a by-product of the LayerNorm backward kernel (xn_out).
//
// Numerics: the two sums are taken over the values AS STORED (bf16-rounded), per 64-column block in fp32 and across
// blocks in fp64; var = E[x^2] - mu^2 is evaluated in fp64 (tests cover a common channel offset of 10 sigma).
#include "common.cuh"
#include "nrvit_internal.h"

namespace nrv {

// one warp per row: stats[row] = (sum x, sum x^2), overwriting
template <typename T>
__global__ void __launch_bounds__(256) rowstats_kernel(const T* __restrict__ x, long long rows, int dim, double* __restrict__ stats) {
  pdl_launch_dependents();   // the next kernel in the stream (a PDL-launched GEMM) may start its prologue
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * 8;
  for (long long r = warp; r < rows; r += nwarps) {
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane * 8; c < dim; c += 256) {
      float v[8];
      V8<T>::load(x + r * dim + c, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s1 += v[j]; s2 = fmaf(v[j], v[j], s2); }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) *reinterpret_cast<double2*>(stats + 2 * r) = make_double2((double)s1, (double)s2);
  }
}

struct FoldJob {
  const void* W;      // [rows, ldw] compute dtype
  void* Wf;           // [rows, ldw] compute dtype: gamma o W
  const float* gamma; // [dim]
  const float* beta;  // [dim]
  const float* bias;  // [rows] or nullptr
  float* c;           // [rows]
  int rows;
};
constexpr int FOLD_MAX_JOBS = 64;   // 2 matrices x 32 layers (ViT-H/14)
struct FoldJobs {
  FoldJob job[FOLD_MAX_JOBS];
  int row_end[FOLD_MAX_JOBS];   // exclusive prefix sums of rows
  int njobs, dim, ldw;
};

// one warp per output row n of one job: Wf[n,:] = gamma o W[n,:] - its mean (rounded to T, zero-sum fix-up),
// c_n = sum_k beta_k W[n,k] + bias_n.  Three passes over a row that sits in L1 / registers.
template <typename T>
__global__ void __launch_bounds__(256) ln_fold_kernel(const __grid_constant__ FoldJobs J) {
  pdl_launch_dependents();   // the next kernel in the stream (a PDL-launched GEMM) may start its prologue
  const int lane = threadIdx.x & 31;
  const int total = J.row_end[J.njobs - 1];
  for (int gr = blockIdx.x * 8 + (threadIdx.x >> 5); gr < total; gr += gridDim.x * 8) {
    int j = 0;
    while (gr >= J.row_end[j]) ++j;
    const FoldJob& f = J.job[j];
    const int n = gr - (j == 0 ? 0 : J.row_end[j - 1]);
    const T* w = reinterpret_cast<const T*>(f.W) + (long long)n * J.ldw;
    T* wf = reinterpret_cast<T*>(f.Wf) + (long long)n * J.ldw;
    // pass 1: row mean of gamma o W, and c
    float s = 0.f, c = 0.f;
    for (int k = lane * 8; k < J.dim; k += 256) {
      float v[8], g[8], b[8];
      V8<T>::load(w + k, v);
      V8<float>::load(f.gamma + k, g);
      V8<float>::load(f.beta + k, b);
#pragma unroll
      for (int e = 0; e < 8; ++e) { s = fmaf(v[e], g[e], s); c = fmaf(v[e], b[e], c); }
    }
    s = warp_sum(s);
    c = warp_sum(c);
    const float m = s / (float)J.dim;
    // pass 2: centred row, rounded with error diffusion: every lane carries the rounding remainder of one element into its
    // next one.  Plain round-to-nearest is NOT enough: W arrives on the bf16 grid, so W - m misses the grid by the same
    // amount for every element of a binade and the errors add coherently (measured: |sum of the rounded row| ~2e-2, a
    // whole weight; with the carry ~2e-4).  Also tracked: the element of smallest magnitude (finest rounding grid).
    float eps_sum = 0.f, best = INFINITY, carry = 0.f;
    int best_k = 0;
    for (int k = lane * 8; k < J.dim; k += 256) {
      float v[8], g[8], o[8];
      V8<T>::load(w + k, v);
      V8<float>::load(f.gamma + k, g);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float t = fmaf(v[e], g[e], -m) + carry;
        o[e] = V8<T>::round(t);
        carry = t - o[e];
        eps_sum += o[e];
        if (fabsf(o[e]) < best) { best = fabsf(o[e]); best_k = k + e; }
      }
      V8<T>::store(wf + k, o);
    }
    eps_sum = warp_sum(eps_sum);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int ok = __shfl_xor_sync(0xffffffffu, best_k, o);
      if (ob < best || (ob == best && ok < best_k)) { best = ob; best_k = ok; }
    }
    __syncwarp();
    // pass 3: the smallest element absorbs the remainder (its new magnitude ~|remainder|, so the re-rounding error is
    // ~2^-9 of the remainder instead of the remainder itself)
    if (lane == 0) {
      wf[best_k] = from_f32<T>(to_f32(wf[best_k]) - eps_sum);
      f.c[n] = c + (f.bias != nullptr ? f.bias[n] : 0.f);
    }
  }
}

int rowstats(const void* x, long long rows, int dim, int dtype, double* stats, cudaStream_t st) {
  if (rows <= 0) return NRV_OK;
  const long long blocks_needed = (rows + 7) / 8;
  const long long cap = (long long)(num_sms() > 0 ? num_sms() : 148) * 8;
  const unsigned grid = (unsigned)(blocks_needed < cap ? blocks_needed : cap);
  if (dtype == NRV_BF16) rowstats_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, rows, dim, stats);
  else rowstats_kernel<float><<<grid, 256, 0, st>>>((const float*)x, rows, dim, stats);
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

// all (layer, matrix) pairs of a forward pass in ONE launch
int ln_fold_weights(const LnFoldJob* jobs, int njobs, int dim, int ldw, int dtype, cudaStream_t st) {
  NRV_REQUIRE(njobs > 0 && njobs <= FOLD_MAX_JOBS, "ln_fold_weights: 1..%d jobs per launch (got %d)", FOLD_MAX_JOBS, njobs);
  NRV_REQUIRE(dim % 8 == 0 && ldw % 8 == 0, "ln_fold_weights: dim and ldw must be multiples of 8");
  FoldJobs J;
  memset(&J, 0, sizeof(J));
  int acc = 0;
  for (int i = 0; i < njobs; ++i) {
    NRV_REQUIRE(jobs[i].W && jobs[i].Wf && jobs[i].gamma && jobs[i].beta && jobs[i].c && jobs[i].rows > 0,
                "ln_fold_weights: null pointer in job %d", i);
    J.job[i].W = jobs[i].W; J.job[i].Wf = jobs[i].Wf; J.job[i].gamma = jobs[i].gamma; J.job[i].beta = jobs[i].beta;
    J.job[i].bias = jobs[i].bias; J.job[i].c = jobs[i].c; J.job[i].rows = jobs[i].rows;
    acc += jobs[i].rows;
    J.row_end[i] = acc;
  }
  J.njobs = njobs; J.dim = dim; J.ldw = ldw;
  const int sms = num_sms() > 0 ? num_sms() : 148;
  const int blocks_needed = (acc + 7) / 8;
  const int grid = blocks_needed < sms * 8 ? blocks_needed : sms * 8;
  if (dtype == NRV_BF16) ln_fold_kernel<bf16><<<grid, 256, 0, st>>>(J);
  else ln_fold_kernel<float><<<grid, 256, 0, st>>>(J);
  count_launch();
  NRV_CUDA(cudaGetLastError());
  return NRV_OK;
}

}  // namespace nrv

using namespace nrv;

extern "C" {

int nrv_rowstats(const void* x, long long rows, int dim, int dtype, double* stats, void* stream) {
  int rc = require_init();
  if (rc) return rc;
  NRV_REQUIRE(dtype == NRV_BF16 || dtype == NRV_F32, "nrv_rowstats: dtype must be NRV_BF16 or NRV_F32");
  NRV_REQUIRE(x && stats, "nrv_rowstats: null pointer");
  NRV_REQUIRE(dim > 0 && dim % 8 == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)stats % 16) == 0,
              "nrv_rowstats: dim must be a multiple of 8, x and stats 16-byte aligned");
  return rowstats(x, rows, dim, dtype, stats, (cudaStream_t)stream);
}

int nrv_ln_fold_weights(const void* W, const float* gamma, const float* beta, const float* bias, void* Wf, float* c,
                        int rows, int dim, long long ldw, int dtype, void* stream) {
  int rc = require_init();
  if (rc) return rc;
  NRV_REQUIRE(dtype == NRV_BF16 || dtype == NRV_F32, "nrv_ln_fold_weights: dtype must be NRV_BF16 or NRV_F32");
  NRV_REQUIRE(((uintptr_t)W % 16) == 0 && ((uintptr_t)Wf % 16) == 0 && ((uintptr_t)gamma % 16) == 0 && ((uintptr_t)beta % 16) == 0,
              "nrv_ln_fold_weights: W, Wf, gamma, beta must be 16-byte aligned");
  NRV_REQUIRE(W && Wf && gamma && beta && c, "nrv_ln_fold_weights: null pointer");
  LnFoldJob j{W, Wf, gamma, beta, bias, c, rows};
  return ln_fold_weights(&j, 1, dim, (int)ldw, dtype, (cudaStream_t)stream);
}

}  // extern "C"
