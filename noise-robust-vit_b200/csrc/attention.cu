// nrv_attn_fwd / nrv_attn_bwd: dispatch between the tcgen05 kernel (attention_tc.cu, production
// bf16 path) and the fp32 CUDA-core kernel (attention_simt.cu, check mode / cross-check).
#include "common.cuh"
#include "nrvit_internal.h"

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
                 int mode, int dtype, int impl, void* stream) {
  int rc = require_init();
  if (rc) return rc;
  rc = attn_common_checks("nrv_attn_fwd", B, N, H, dh, mode, dtype, impl);
  if (rc) return rc;
  NRV_REQUIRE(qkv && out, "nrv_attn_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == NRV_ATTN_SINKHORN3) return sinkhorn_fwd(qkv, out, lse, B, N, H, dh, scale, dtype, st);
  const bool tc_ok = attn_tc_supported(N, dh, dtype);
  const bool big_ok = attn_big_supported(N, dh, dtype);     // general tcgen05 forward (dh <= 128, N <= 384)
  if (impl == NRV_ATTN_IMPL_TC && !tc_ok && !big_ok) {
    set_error("nrv_attn_fwd: tcgen05 attention does not support N=%d dh=%d dtype=%d", N, dh, dtype);
    return NRV_ENOTIMPL;
  }
  if (impl != NRV_ATTN_IMPL_SIMT) {
    if (tc_ok) return attn_fwd_tc(qkv, out, lse, B, N, H, dh, scale, st);
    if (big_ok) return attn_fwd_big(qkv, out, lse, B, N, H, dh, scale, st);
  }
  return attn_fwd_simt(qkv, out, lse, B, N, H, dh, scale, dtype, st);
}

int nrv_attn_debug_timestamps(long long* device_buf) { attn_tc_set_debug(device_buf); return NRV_OK; }

size_t nrv_attn_bwd_workspace(int B, int N, int H) {
  const size_t delta = (size_t)B * N * H * sizeof(float) + 256;          // softmax: rowsum(dO o O)
  const size_t sk = sinkhorn_bwd_scratch_bytes(B, N, H);                 // Sinkhorn: per-CTA N x N gradient matrix
  return delta > sk ? delta : sk;
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
  if (mode == NRV_ATTN_SINKHORN3) {
    NRV_REQUIRE(workspace != nullptr && workspace_bytes >= nrv_attn_bwd_workspace(B, N, H),
                "nrv_attn_bwd: workspace of nrv_attn_bwd_workspace() bytes required");
    return sinkhorn_bwd(qkv, dout, lse, dqkv, (float*)workspace, B, N, H, dh, scale, dtype, st);
  }
  const bool tc_ok = attn_tc_supported(N, dh, dtype);
  const bool bwd2_ok = attn_bwd2_supported(N, dh, dtype);   // fused backward: dh = 64, up to 256 tokens
  if (impl == NRV_ATTN_IMPL_TC && !tc_ok && !bwd2_ok) {
    set_error("nrv_attn_bwd: tcgen05 attention does not support N=%d dh=%d dtype=%d", N, dh, dtype);
    return NRV_ENOTIMPL;
  }
  if (impl != NRV_ATTN_IMPL_SIMT && tc_ok)
  {
    NRV_REQUIRE(workspace != nullptr && workspace_bytes >= nrv_attn_bwd_workspace(B, N, H),
                "nrv_attn_bwd: workspace of nrv_attn_bwd_workspace() bytes required");
    return attn_bwd_tc(qkv, out, dout, lse, dqkv, (float*)workspace, B, N, H, dh, scale, st);
  }
  if (impl != NRV_ATTN_IMPL_SIMT && bwd2_ok) return attn_bwd_tc2(qkv, out, dout, lse, dqkv, B, N, H, dh, scale, st);
  return attn_bwd_simt(qkv, out, dout, lse, dqkv, B, N, H, dh, scale, dtype, st);
}

}  // extern "C"
