// nrv_attn_fwd / nrv_attn_bwd: dispatch between the tcgen05 kernels (attention_fwd2.cu / attention_bwd2.cu for
// dh = 64 and up to 208 / 256 tokens, attention_fwd_big.cu / attention_bwd_big.cu for the general shapes: production bf16 path), the
// Sinkhorn kernels (attention_sinkhorn.cu) and the fp32 CUDA-core kernels (attention_simt.cu, check mode / cross-check).
#include "common.cuh"
#include "nrvit_internal.h"

namespace nrv {
static long long* g_attn_dbg = nullptr;   // optional device buffer for the phase timestamps of the tcgen05 kernels
void attn_tc_set_debug(long long* buf) { g_attn_dbg = buf; }
long long* attn_tc_get_debug() { return g_attn_dbg; }
// shapes of the two-pipeline training forward (attention_fwd2.cu): one 64-wide head, all keys in one TMEM tile
bool attn_tc_supported(int N, int dh, int dtype) { return dtype == NRV_BF16 && dh == 64 && N >= 1 && N <= 208; }
}  // namespace nrv

using namespace nrv;

static int attn_common_checks(const char* who, int B, int N, int H, int dh, int mode, int dtype, int impl) {
  NRV_REQUIRE(dtype == NRV_BF16 || dtype == NRV_F32, "%s: dtype must be NRV_BF16 or NRV_F32", who);
  NRV_REQUIRE(B > 0 && N > 0 && H > 0 && dh > 0, "%s: B, N, H, dh must be positive", who);
  NRV_REQUIRE(impl >= NRV_ATTN_IMPL_AUTO && impl <= NRV_ATTN_IMPL_TC, "%s: bad impl %d", who, impl);
  NRV_REQUIRE(mode == NRV_ATTN_SOFTMAX || mode == NRV_ATTN_SINKHORN3, "%s: bad attention mode %d", who, mode);
  return NRV_OK;
}

extern "C" {

int nrv_attn_fwd(const void* qkv, void* out, float* lse, int B, int N, int H, int dh, float scale,
                 int mode, int dtype, int impl, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = require_init();
  if (rc) return rc;
  rc = attn_common_checks("nrv_attn_fwd", B, N, H, dh, mode, dtype, impl);
  if (rc) return rc;
  NRV_REQUIRE(qkv && out, "nrv_attn_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == NRV_ATTN_SINKHORN3) {
    // tensor-core kernel for bf16, dh = 64, up to 208 tokens; the CUDA-core kernel otherwise (same statistics layout)
    if (impl != NRV_ATTN_IMPL_SIMT && sinkhorn_tc_supported(N, dh, dtype)) return sinkhorn_fwd_tc(qkv, out, lse, B, N, H, dh, scale, st);
    if (impl == NRV_ATTN_IMPL_TC) {
      set_error("nrv_attn_fwd: tcgen05 Sinkhorn attention does not support N=%d dh=%d dtype=%d", N, dh, dtype);
      return NRV_ENOTIMPL;
    }
    return sinkhorn_fwd(qkv, out, lse, workspace, workspace_bytes, B, N, H, dh, scale, dtype, st);
  }
  const bool tc_ok = attn_tc_supported(N, dh, dtype);
  const bool big_ok = attn_big_supported(N, dh, dtype);     // general tcgen05 forward (dh <= 128, N <= 384)
  if (impl == NRV_ATTN_IMPL_TC && !tc_ok && !big_ok) {
    set_error("nrv_attn_fwd: tcgen05 attention does not support N=%d dh=%d dtype=%d", N, dh, dtype);
    return NRV_ENOTIMPL;
  }
  if (impl != NRV_ATTN_IMPL_SIMT) {
    if (tc_ok) return attn_fwd_tc2(qkv, out, lse, B, N, H, dh, scale, st);
    if (big_ok) return attn_fwd_big(qkv, out, lse, B, N, H, dh, scale, st);
  }
  return attn_fwd_simt(qkv, out, lse, B, N, H, dh, scale, dtype, st);
}

int nrv_attn_debug_timestamps(long long* device_buf) { attn_tc_set_debug(device_buf); return NRV_OK; }

size_t nrv_attn_bwd_workspace(int B, int N, int H, int dh) {
  const size_t delta = (size_t)B * N * H * sizeof(float) + 256;          // softmax: rowsum(dO o O)
  const size_t sk = sinkhorn_bwd_scratch_bytes(B, N, H, dh);             // Sinkhorn: per-CTA N x N gradient (+ probability) matrix
  size_t big = 0;                                                        // general tcgen05 backward: per-CTA running dQ
  // (also for the shapes of the fused backward: attention dropout runs the general kernel on every shape it supports)
  if (attn_bwd_big_supported(N, dh, NRV_BF16)) big = attn_bwd_big_scratch_bytes(B, N, H, dh);
  const size_t m = delta > sk ? delta : sk;
  return m > big ? m : big;
}

size_t nrv_attn_fwd_workspace(int B, int N, int H, int dh, int mode) {
  return mode == NRV_ATTN_SINKHORN3 ? sinkhorn_fwd_scratch_bytes(B, N, H, dh) : 0;
}

int nrv_attn_probs(const void* qkv, float* probs, float* stats, int B, int N, int H, int dh, float scale, int mode,
                   int dtype, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = require_init();
  if (rc) return rc;
  rc = attn_common_checks("nrv_attn_probs", B, N, H, dh, mode, dtype, NRV_ATTN_IMPL_AUTO);
  if (rc) return rc;
  NRV_REQUIRE(qkv && probs && stats, "nrv_attn_probs: null pointer");
  return attn_probs(qkv, probs, stats, workspace, workspace_bytes, B, N, H, dh, scale, mode == NRV_ATTN_SINKHORN3, dtype,
                    (cudaStream_t)stream);
}

size_t nrv_attn_stats_elems(int B, int N, int H, int mode) {
  return (size_t)B * H * N * (mode == NRV_ATTN_SINKHORN3 ? 8 : 1);
}

int nrv_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                 int B, int N, int H, int dh, float scale, int mode, int dtype, int impl, void* workspace,
                 size_t workspace_bytes, void* stream) {
  int rc = require_init();
  if (rc) return rc;
  rc = attn_common_checks("nrv_attn_bwd", B, N, H, dh, mode, dtype, impl);
  if (rc) return rc;
  NRV_REQUIRE(qkv && out && dout && lse && dqkv, "nrv_attn_bwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == NRV_ATTN_SINKHORN3 && impl != NRV_ATTN_IMPL_SIMT && sinkhorn_tc_supported(N, dh, dtype))
    return sinkhorn_bwd_tc(qkv, out, dout, lse, dqkv, B, N, H, dh, scale, st);
  if (mode == NRV_ATTN_SINKHORN3) {
    if (impl == NRV_ATTN_IMPL_TC) {
      set_error("nrv_attn_bwd: tcgen05 Sinkhorn attention does not support N=%d dh=%d dtype=%d", N, dh, dtype);
      return NRV_ENOTIMPL;
    }
    NRV_REQUIRE(workspace != nullptr && workspace_bytes >= nrv_attn_bwd_workspace(B, N, H, dh),
                "nrv_attn_bwd: workspace of nrv_attn_bwd_workspace() bytes required");
    return sinkhorn_bwd(qkv, dout, lse, dqkv, (float*)workspace, B, N, H, dh, scale, dtype, st);
  }
  const bool bwd2_ok = attn_bwd2_supported(N, dh, dtype);   // fused backward: dh = 64, up to 256 tokens
  const bool big_ok = attn_bwd_big_supported(N, dh, dtype); // general backward: dh <= 80, up to 1024 tokens
  if (impl == NRV_ATTN_IMPL_TC && !bwd2_ok && !big_ok) {
    set_error("nrv_attn_bwd: tcgen05 attention does not support N=%d dh=%d dtype=%d", N, dh, dtype);
    return NRV_ENOTIMPL;
  }
  if (impl != NRV_ATTN_IMPL_SIMT && bwd2_ok) return attn_bwd_tc2(qkv, out, dout, lse, dqkv, B, N, H, dh, scale, st);
  if (impl != NRV_ATTN_IMPL_SIMT && big_ok)
    return attn_bwd_big(qkv, out, dout, lse, dqkv, (float*)workspace, workspace_bytes, B, N, H, dh, scale, st);
  return attn_bwd_simt(qkv, out, dout, lse, dqkv, B, N, H, dh, scale, dtype, st);
}

}  // extern "C"
