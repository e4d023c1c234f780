// tcgen05 attention (placeholder until the kernel lands): reports "unsupported" so the dispatcher
// never selects it.
#include "common.cuh"
#include "nrvit_internal.h"

namespace nrv {

bool attn_tc_supported(int N, int dh, int dtype) { (void)N; (void)dh; (void)dtype; return false; }

int attn_fwd_tc(const void*, void*, float*, int, int, int, int, float, cudaStream_t) {
  set_error("tcgen05 attention forward not built");
  return NRV_ENOTIMPL;
}
int attn_bwd_tc(const void*, const void*, const void*, const float*, void*, int, int, int, int, float, cudaStream_t) {
  set_error("tcgen05 attention backward not built");
  return NRV_ENOTIMPL;
}

}  // namespace nrv
