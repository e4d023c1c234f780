// Runtime entry points of the C ABI: init, error reporting, TMA descriptor encoding.
#include "common.cuh"
#include "nrvit_internal.h"

#include <atomic>
#include <stdarg.h>
#include <string.h>

namespace nrv {

static thread_local char g_err[512] = "";
static int g_device = -1;
static int g_sms = 0;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return NRV_ECUDA;
}

bool initialised() { return g_device >= 0; }
int require_init() {
  if (g_device < 0) {
    set_error("libnrvit not initialised: call nrv_init(device) on an sm_100 GPU (no CPU fallback)");
    return NRV_ENOTINIT;
  }
  return NRV_OK;
}
int num_sms() { return g_sms; }

int encode_tmap_2d(CUtensorMap* map, CUtensorMapDataType dt, const void* gptr, uint64_t inner,
                   uint64_t outer, uint64_t outer_stride_bytes, uint32_t box_inner,
                   uint32_t box_outer, CUtensorMapSwizzle swz) {
  if (!g_encode) return require_init();
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {outer_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, dt, 2, const_cast<void*>(gptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(2d) failed: CUresult %d (ptr %p dims %llu x %llu stride %llu box %u x %u)",
              (int)r, gptr, (unsigned long long)inner, (unsigned long long)outer,
              (unsigned long long)outer_stride_bytes, box_inner, box_outer);
    return NRV_ECUDA;
  }
  return NRV_OK;
}

int encode_tmap_3d(CUtensorMap* map, CUtensorMapDataType dt, const void* gptr, uint64_t d0,
                   uint64_t d1, uint64_t d2, uint64_t stride1_bytes, uint64_t stride2_bytes,
                   uint32_t b0, uint32_t b1, uint32_t b2, CUtensorMapSwizzle swz) {
  if (!g_encode) return require_init();
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(map, dt, 3, const_cast<void*>(gptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d) failed: CUresult %d (ptr %p dims %llu,%llu,%llu strides %llu,%llu box %u,%u,%u)",
              (int)r, gptr, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
              (unsigned long long)stride1_bytes, (unsigned long long)stride2_bytes, b0, b1, b2);
    return NRV_ECUDA;
  }
  return NRV_OK;
}

int encode_tmap_4d(CUtensorMap* map, CUtensorMapDataType dt, const void* gptr, const uint64_t (&dims_)[4],
                   const uint64_t (&strides_bytes)[3], const uint32_t (&box_)[4], CUtensorMapSwizzle swz) {
  if (!g_encode) return require_init();
  cuuint64_t dims[4] = {dims_[0], dims_[1], dims_[2], dims_[3]};
  cuuint64_t strides[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
  cuuint32_t box[4] = {box_[0], box_[1], box_[2], box_[3]};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(map, dt, 4, const_cast<void*>(gptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(4d) failed: CUresult %d (ptr %p dims %llu,%llu,%llu,%llu box %u,%u,%u,%u)", (int)r, gptr,
              (unsigned long long)dims_[0], (unsigned long long)dims_[1], (unsigned long long)dims_[2],
              (unsigned long long)dims_[3], box_[0], box_[1], box_[2], box_[3]);
    return NRV_ECUDA;
  }
  return NRV_OK;
}

int encode_tmap_5d(CUtensorMap* map, CUtensorMapDataType dt, const void* gptr, const uint64_t (&dims_)[5],
                   const uint64_t (&strides_bytes)[4], const uint32_t (&box_)[5], CUtensorMapSwizzle swz) {
  if (!g_encode) return require_init();
  cuuint64_t dims[5] = {dims_[0], dims_[1], dims_[2], dims_[3], dims_[4]};
  cuuint64_t strides[4] = {strides_bytes[0], strides_bytes[1], strides_bytes[2], strides_bytes[3]};
  cuuint32_t box[5] = {box_[0], box_[1], box_[2], box_[3], box_[4]};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = g_encode(map, dt, 5, const_cast<void*>(gptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(5d) failed: CUresult %d (ptr %p dims %llu,%llu,%llu,%llu,%llu box %u,%u,%u,%u,%u)", (int)r, gptr,
              (unsigned long long)dims_[0], (unsigned long long)dims_[1], (unsigned long long)dims_[2],
              (unsigned long long)dims_[3], (unsigned long long)dims_[4], box_[0], box_[1], box_[2], box_[3], box_[4]);
    return NRV_ECUDA;
  }
  return NRV_OK;
}

}  // namespace nrv

using namespace nrv;

extern "C" {

int nrv_abi_version(void) { return NRV_ABI_VERSION; }

const char* nrv_last_error(void) { return g_err; }

int nrv_num_sms(void) { return g_sms; }

long long nrv_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int nrv_init(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    set_error("nrv_init: no CUDA device visible (%s); libnrvit has no CPU fallback",
              e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
    return NRV_EARCH;
  }
  NRV_REQUIRE(device >= 0 && device < count, "nrv_init: device %d out of range [0,%d)", device, count);
  cudaDeviceProp prop;
  NRV_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("nrv_init: device %d is sm_%d%d; libnrvit is built for sm_100a only", device, prop.major,
              prop.minor);
    return NRV_EARCH;
  }
  NRV_CUDA(cudaSetDevice(device));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  NRV_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    set_error("nrv_init: cuTensorMapEncodeTiled not available from the driver");
    return NRV_ECUDA;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  g_sms = prop.multiProcessorCount;
  g_device = device;
  return NRV_OK;
}

int nrv_gemm(const nrv_gemm_desc* d, void* stream) {
  int rc = require_init();
  if (rc) return rc;
  return gemm_dispatch(d, reinterpret_cast<cudaStream_t>(stream));
}

int nrv_gemm_timing(int enable) { gemm_timing_enable(enable); return NRV_OK; }
int nrv_gemm_timing_detail(long long* out, int max_records) { return gemm_timing_detail(out, max_records); }
int nrv_gemm_timing_read(double* ms, double* flops, long long* launches) { return gemm_timing_read(ms, flops, launches); }

size_t nrv_gemm_workspace_bytes(int M, int N, int K, int dtype) { return gemm_workspace_bytes(M, N, K, dtype); }

}  // extern "C"
