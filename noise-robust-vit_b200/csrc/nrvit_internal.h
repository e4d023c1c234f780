// Internal declarations shared by the translation units of libnrvit.
#pragma once
#include "../../include/nrvit.h"
#include <cuda_runtime.h>

namespace nrv {
int gemm_dispatch(const nrv_gemm_desc* d, cudaStream_t stream);
bool initialised();
int require_init();
}  // namespace nrv
