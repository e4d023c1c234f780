// Common device/host helpers for libnrvit (sm_100a only).
//
// Thin inline-PTX wrappers around the Blackwell primitives the kernels use:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld / fences),
// plus the shared-memory / instruction descriptor encoders for UMMA.
// No torch headers, no CUTLASS dependency.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/nrvit.h"

#ifndef __CUDA_ARCH_FEAT_SM100_ALL
#if defined(__CUDA_ARCH__)
#error "libnrvit must be compiled with -gencode arch=compute_100a,code=sm_100a"
#endif
#endif

namespace nrv {

typedef __nv_bfloat16 bf16;

// ----------------------------------------------------------------------------------------------
// error plumbing (host)
// ----------------------------------------------------------------------------------------------

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define NRV_CUDA(call)                                                       \
  do {                                                                       \
    cudaError_t _e = (call);                                                 \
    if (_e != cudaSuccess) return ::nrv::cuda_fail(_e, #call, __FILE__, __LINE__); \
  } while (0)

#define NRV_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      ::nrv::set_error(__VA_ARGS__);           \
      return NRV_EINVAL;                \
    }                                          \
  } while (0)

// TMA descriptor encode (driver entry point resolved at nrv_init; no libcuda link dependency)
int encode_tmap_2d(CUtensorMap* map, CUtensorMapDataType dt, const void* gptr, uint64_t inner,
                   uint64_t outer, uint64_t outer_stride_bytes, uint32_t box_inner,
                   uint32_t box_outer, CUtensorMapSwizzle swz);
int encode_tmap_3d(CUtensorMap* map, CUtensorMapDataType dt, const void* gptr, uint64_t d0,
                   uint64_t d1, uint64_t d2, uint64_t stride1_bytes, uint64_t stride2_bytes,
                   uint32_t b0, uint32_t b1, uint32_t b2, CUtensorMapSwizzle swz);
int encode_tmap_4d(CUtensorMap* map, CUtensorMapDataType dt, const void* gptr, const uint64_t (&dims)[4],
                   const uint64_t (&strides_bytes)[3], const uint32_t (&box)[4], CUtensorMapSwizzle swz);
int encode_tmap_5d(CUtensorMap* map, CUtensorMapDataType dt, const void* gptr, const uint64_t (&dims)[5],
                   const uint64_t (&strides_bytes)[4], const uint32_t (&box)[5], CUtensorMapSwizzle swz);
int num_sms();
void count_launch(int n = 1);

// ----------------------------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------------------------
#if defined(__CUDACC__)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier --------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(0x989680u)   // suspend-time hint: the warp sleeps in hardware until the phase
      : "memory");                              // completes instead of spinning through the issue slots
  return ok != 0;
}
// Watchdog: a lost arrive must become a trap (launch failure), never a hung GPU.
#ifndef NRV_WATCHDOG_CYCLES
#define NRV_WATCHDOG_CYCLES 6000000000ll  // ~3 s at 1.9 GHz
#endif
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity, int tag) {
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > NRV_WATCHDOG_CYCLES) {
      printf("[nrvit] mbarrier watchdog: block %d thread %d tag %d parity %u\n", blockIdx.x,
             threadIdx.x, tag, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag = 0) {
#pragma unroll 1
  for (int i = 0; i < 64; ++i)
    if (mbar_try_wait(bar, parity)) return;
  mbar_wait_slow(bar, parity, tag);
}

// Call-free variant for code that holds a large register working set across the wait (a call to the
// noinline watchdog would spill it): bounded spin, then trap.
__device__ __forceinline__ void mbar_wait_inline(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its predecessor in the stream
// is still draining; pdl_wait() blocks until that predecessor has completed and its memory is visible, so everything
// before it (barrier init, TMEM allocation, descriptor prefetch) overlaps the predecessor's tail.  Without the launch
// attribute it returns at once.  pdl_launch_dependents() lets the NEXT kernel in the stream begin that early start.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- explicit shared-memory vector access (a generic pointer into dynamic smem compiles to LD.E / ST.E) ------
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts128f(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// ---- proxy fences ----------------------------------------------------------------------------
__device__ __forceinline__ void fence_async_smem() {
  // generic-proxy smem writes -> visible to the async proxy (TMA / UMMA operand reads)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- TMA -------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
      "{%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
      "{%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
      "{%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// 1-D bulk copy global -> shared (no tensor map): `bytes` contiguous bytes, multiple of 16, 16-byte aligned on both sides
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
               : "memory");
}

// TMA store smem -> global (bulk async group completion); out-of-bounds elements are clipped
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until the smem source of all but the newest `N` bulk groups has been read
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---- tcgen05 ---------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// whole warp; writes TMEM base address to *smem_dst
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; one thread issues
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread retire
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i gets row (lane base + i), v[j] = column j
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
        "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
        "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (M = 128 lanes, bf16 pairs packed two per 32-bit
// column, K-major) comes straight from tensor memory, e.g. softmax probabilities written by tcgen05.st
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 8 consecutive 32-bit columns, thread i writes row (lane base + i)
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

// ---- packed fp32 pairs (FFMA2 / FADD2 / FMUL2: one issue slot for two lanes of math) -----------
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// ---- order-pinned primitives for the exp phases of the attention kernels ------------------------------
// One warp per scheduler has to keep the MUFU pipe (8 cycles per warp-wide ex2) busy on its own.  A warp issues
// in order, so every consumer placed right behind its ex2 stalls for the MUFU latency (measured: 233 instead of
// 128 cycles per 16 scores).  The stream below is software-pipelined by hand -- FFMA2 on chunk k+1, ex2 on
// chunk k, sum / bf16 pack on chunk k-1, interleaved pair by pair -- and `asm volatile` keeps ptxas from
// re-fusing the stages.  All of them work in place on the b32 registers tcgen05.ld delivered.
__device__ __forceinline__ void pv_fma2(uint32_t& a, uint32_t& b, uint64_t c2, uint64_t n2) {
  asm volatile("{\n\t.reg .b64 t;\n\tmov.b64 t, {%0, %1};\n\tfma.rn.f32x2 t, t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}"
               : "+r"(a), "+r"(b) : "l"(c2), "l"(n2));
}
__device__ __forceinline__ void pv_ex2(uint32_t& a) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(a)); }
__device__ __forceinline__ void pv_add2(uint64_t& s, uint32_t a, uint32_t b) {
  asm volatile("{\n\t.reg .b64 t;\n\tmov.b64 t, {%1, %2};\n\tadd.rn.f32x2 %0, %0, t;\n\t}" : "+l"(s) : "r"(a), "r"(b));
}
__device__ __forceinline__ uint32_t pv_pack(uint32_t lo, uint32_t hi) {
  uint32_t d;
  asm volatile("cvt.rn.bf16x2.f32 %0, %2, %1;" : "=r"(d) : "r"(lo), "r"(hi));
  return d;
}

__device__ __forceinline__ void pv_add2r(uint32_t& a, uint32_t& b, uint64_t c) {      // (a, b) += c
  asm volatile("{\n\t.reg .b64 t;\n\tmov.b64 t, {%0, %1};\n\tadd.rn.f32x2 t, t, %2;\n\tmov.b64 {%0, %1}, t;\n\t}" : "+r"(a), "+r"(b) : "l"(c));
}
__device__ __forceinline__ void pv_mul2r(uint32_t& a, uint32_t& b, uint32_t p0, uint32_t p1) {   // (a, b) *= (p0, p1)
  asm volatile("{\n\t.reg .b64 t, u;\n\tmov.b64 t, {%0, %1};\n\tmov.b64 u, {%2, %3};\n\tmul.rn.f32x2 t, t, u;\n\tmov.b64 {%0, %1}, t;\n\t}"
               : "+r"(a), "+r"(b) : "r"(p0), "r"(p1));
}

// ---- CTA pair (cta_group::2) variants ----------------------------------------------------------
// Two CTAs of a cluster on one TPC execute ONE MMA of M = 256: each CTA stages its 128 rows of A and
// its half of B; the leader (cluster rank 0) issues, accumulator rows split across the two TMEMs.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Arrive on a (possibly remote) CTA's mbarrier.  Default semantics on purpose: `.release.cluster` compiles to
// MEMBAR.ALL.CTA + ERRBAR in front of the arrive (14 % of the GELU GEMM's stall samples, ncu); the only thing this
// arrive publishes is "my tcgen05.ld of the accumulator is done", which tcgen05.wait::ld + tcgen05.fence order.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into THIS CTA's smem whose completion bytes are credited to the leader CTA's mbarrier
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
      "{%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
      "{%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3,
                                                 int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
      "{%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at the same smem offset in every CTA of `mask` once the pair's MMAs retire
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   bar),
               "h"(mask)
               : "memory");
}

// ---- UMMA descriptors ------------------------------------------------------------------------
// 64-bit shared-memory matrix descriptor, SWIZZLE_128B canonical layouts (sm_100 "version 1"):
//   bits [0,14)  start address >> 4       bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4  bits [46,48) version = 1
//   bits [61,64) layout type (2 = SWIZZLE_128B)
// K-major  : rows of 128 B (64 bf16 of K), 8-row groups 1024 B apart (SBO); LBO unused.
// MN-major : rows of 128 B (64 bf16 of M/N), one row per k; 8-k groups SBO apart,
//            64-element M/N chunks LBO apart.
__host__ __device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes,
                                                                  uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Same descriptor with the layout type given: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B (rows / lines of 128, 64, 32
// bytes; the 16-byte units of a line are XORed with the line index modulo 8, 4, 2).  K-major: 8-row groups SBO apart;
// MN-major: groups of (line bytes / 2) M/N elements LBO apart, 8-k groups SBO apart.
__host__ __device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}

// 32-bit instruction descriptor for kind::f16 / kind::tf32 with fp32 accumulate.
//   [4,6) c fmt (1 = f32)   [7,10) a fmt   [10,13) b fmt  (0 f16, 1 bf16, 2 tf32)
//   [15] a major  [16] b major (0 = K, 1 = MN)   [17,23) N>>3   [24,29) M>>4
__host__ __device__ __forceinline__ uint32_t make_idesc(uint32_t fmt, uint32_t a_mn, uint32_t b_mn,
                                                        uint32_t M, uint32_t N) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) |
         ((M >> 4) << 24);
}

// ---- small math ------------------------------------------------------------------------------
__device__ __forceinline__ float exp2f_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// MUFU.RCP alone: __fdividef(1, x) adds a range test, a select and two scalings (5 issue slots instead of 1); every
// use below has 1 <= x < 2^100.
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// bf16x2 product with one rounding per lane (HMUL2.BF16): dX epilogue factor, as a bf16 autocast graph would apply it
__device__ __forceinline__ uint32_t bf2_mul(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
// GELU (exact erf form, nn.GELU() default: simple_vit.py:40 ; vit.py:44) and its derivative on the
// epilogue's instruction budget.  Phi(u) = 0.5 (1 + erf(u / sqrt 2)) through Abramowitz-Stegun 7.1.26
//   erf(z) = 1 - (a1 t + ... + a5 t^5) exp(-z^2),  t = 1 / (1 + p z),  z >= 0,  |error| <= 1.5e-7
// evaluated as the TAIL Q/2 = 0.5 poly(t) exp(-u^2/2), so negative arguments keep full relative
// precision (no 1 - erf cancellation) and exp(-u^2/2) is shared with the density in GELU'.
// ~13 instructions incl. 2 MUFU (rcp, ex2) instead of ~30 for erff + expf.
__device__ __forceinline__ void gelu_parts(float u, float& cdf, float& e) {
  const float au = fabsf(u);
  const float t = rcp_approx(fmaf(au, 0.3275911f * 0.70710678118654752440f, 1.0f));
  e = exp2f_approx(-0.72134752044448170368f * u * u);   // exp(-u^2 / 2)
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(t, poly, 1.421413741f);
  poly = fmaf(t, poly, -0.284496736f);
  poly = fmaf(t, poly, 0.254829592f);
  const float half_q = 0.5f * poly * t * e;               // upper tail of the normal at |u|
  cdf = u >= 0.f ? 1.0f - half_q : half_q;
}
__device__ __forceinline__ float gelu_erf(float x) {
  float cdf, e;
  gelu_parts(x, cdf, e);
  return x * cdf;
}
__device__ __forceinline__ float dgelu_erf(float x) {
  float cdf, e;
  gelu_parts(x, cdf, e);
  return fmaf(x * 0.39894228040143267794f, e, cdf);
}
// Fast GELU pair for the bf16 production epilogues (two lanes of packed f32x2 math per issue slot).
//   Phi(u) = 0.5 + 0.5 tanh(y(u)),  y(u) = atanh(erf(u / sqrt 2)) = u * P(u^2)   (odd, smooth; the familiar "tanh GELU" is
//   the two-term truncation of y with hand-picked constants).  P is a degree-3 minimax fit on u^2 <= 36 (u^2 clamped
//   beyond; tools/fit_gelu.py): max |GELU error| 1.9e-5, max |GELU' error| 6.3e-5 over all u (degree 2: 4.3e-5 / 1.3e-4,
//   which showed next to bf16 rounding in the negative tail where GELU' is small).  GELU' is the exact derivative of the approximation:
//   Phi + u (1 - t^2) y'(u) / 2,  y' = sum (2k+1) c_k u^2k.  One MUFU.TANH per element: the earlier sigmoid form
//   (EX2 -> +1 -> RCP, degree-4 P) left two dependent MUFU round trips per element on warps that have only one
//   partner per scheduler; measured on the FC1 GEMM of ViT-B/16 (50432 x 3072 x 768, B200): GELU + GELU' epilogue
//   246.6 -> 225.1 us, GELU alone 203.0 -> 195.3 us, plain store 189 us; error against the exact erf form on fp64
//   equal to bf16 rounding of the exact value to three digits (tools/gpu_time_gelu_gemm.py).
//   19 issue slots per two elements for h and g (was 24).  The fp32 check mode keeps the Abramowitz-Stegun form above.
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void gelu_sig_pair(float u0, float u1, bool want_grad, float& h0, float& h1, float& g0, float& g1) {
  constexpr float c0 = 7.9780044257e-01f, c1 = 3.6594009760e-02f, c2 = -2.1684620180e-04f, c3 = -1.1194026557e-05f;
  const uint64_t u2 = f2_pack(u0, u1);
  float s0, s1;
  f2_unpack(f2_mul(u2, u2), s0, s1);
  const uint64_t s2 = f2_pack(fminf(s0, 36.f), fminf(s1, 36.f));
  uint64_t p2 = f2_fma(s2, f2_pack(c3, c3), f2_pack(c2, c2));
  p2 = f2_fma(p2, s2, f2_pack(c1, c1));
  p2 = f2_fma(p2, s2, f2_pack(c0, c0));
  float y0, y1;
  f2_unpack(f2_mul(u2, p2), y0, y1);
  const uint64_t t2 = f2_pack(tanh_approx(y0), tanh_approx(y1));
  const uint64_t cdf2 = f2_fma(t2, f2_pack(0.5f, 0.5f), f2_pack(0.5f, 0.5f));
  f2_unpack(f2_mul(u2, cdf2), h0, h1);
  if (want_grad) {
    uint64_t q2 = f2_fma(s2, f2_pack(3.5f * c3, 3.5f * c3), f2_pack(2.5f * c2, 2.5f * c2));   // y'(u) / 2 = sum (2k+1) c_k s^k / 2
    q2 = f2_fma(q2, s2, f2_pack(1.5f * c1, 1.5f * c1));
    q2 = f2_fma(q2, s2, f2_pack(0.5f * c0, 0.5f * c0));
    const uint64_t om2 = f2_fma(t2, f2_mul(t2, f2_pack(-1.f, -1.f)), f2_pack(1.f, 1.f));        // 1 - t^2
    f2_unpack(f2_fma(u2, f2_mul(om2, q2), cdf2), g0, g1);
  }
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(t);
}
// ---- dropout on the attention probabilities (vit.py:105-110 ; README ViT Attention.dropout) -------------------------
// keep_ij is the counter-based decision of nrv_dropout for element e = ((b*H + h)*N + i)*N + j at site
// NRV_DROP_ATTN_PROB: Philox4x32-10 call e / 4 (key = seed, counter word 2 = stream id), component e % 4, kept when
// (bits >> 8) >= p * 2^24.  Shared by the CUDA-core kernels (attention_simt.cu) and the general tcgen05 kernels.
struct AttnDrop {
  float p;              // 0 = no dropout
  uint2 key;            // seed
  uint32_t stream_id;   // (layer + 1) * 8 + site, as in nrv_dropout
};
static inline AttnDrop attn_make_drop(float p, unsigned long long seed, int layer) {
  AttnDrop dr;
  dr.p = p;
  dr.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  dr.stream_id = (uint32_t)(layer + 1) * 8u + 4u;   // site NRV_DROP_ATTN_PROB (include/nrvit.h)
  return dr;
}
// keep bits of Philox call q: bit c set = element 4q + c survives
__device__ __forceinline__ uint32_t attn_keep4(const AttnDrop& dr, uint32_t thresh, unsigned long long q) {
  uint4 c = make_uint4((uint32_t)q, (uint32_t)(q >> 32), dr.stream_id, 0u);
  uint2 k = dr.key;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
  }
  return ((c.x >> 8) >= thresh ? 1u : 0u) | ((c.y >> 8) >= thresh ? 2u : 0u) | ((c.z >> 8) >= thresh ? 4u : 0u) |
         ((c.w >> 8) >= thresh ? 8u : 0u);
}
// keep bits of the 16 consecutive elements e0 .. e0+15 (bit t = element e0 + t): 4 calls when e0 is a multiple of 4, else 5
__device__ __forceinline__ uint32_t attn_keep16(const AttnDrop& dr, uint32_t thresh, unsigned long long e0) {
  const unsigned long long q0 = e0 >> 2;
  const uint32_t sh = (uint32_t)e0 & 3u;
  uint32_t bits = 0;
#pragma unroll
  for (int g = 0; g < 4; ++g) bits |= attn_keep4(dr, thresh, q0 + g) << (4 * g);
  if (sh != 0) bits |= attn_keep4(dr, thresh, q0 + 4) << 16;
  return (bits >> sh) & 0xffffu;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- softmax building blocks of the tcgen05 attention forward kernels ------------------------------------
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// row max over one 16-column chunk (columns c0 .. c0+15; only the row's last chunk can hold columns >= N)
__device__ __forceinline__ void f2_max16(const uint32_t (&v)[16], int c0, int N, float& m0, float& m1) {
  if (c0 + 16 <= N) {
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      m0 = fmax3(m0, __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
      m1 = fmax3(m1, __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c0 + j < N) m0 = fmaxf(m0, __uint_as_float(v[j]));
  }
}
// p = exp2(s * c - mx * c) for chunk c, accumulated into the packed row sum, written as bf16 pairs to TMEM
__device__ __forceinline__ void f2_exp16(const uint32_t (&v)[16], int c, int N, uint64_t c2, uint64_t noff2,
                                         uint64_t& sum2, uint32_t t_p) {
  uint32_t pk[8];
  const int c0 = c * 16;
  const bool full = c0 + 16 <= N;
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    const uint64_t x2 = f2_fma(f2_pack(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), c2, noff2);
    float x0, x1;
    f2_unpack(x2, x0, x1);
    float e0 = ex2f(x0), e1 = ex2f(x1);
    if (!full) {
      if (c0 + j >= N) e0 = 0.f;
      if (c0 + j + 1 >= N) e1 = 0.f;
    }
    sum2 = f2_add(sum2, f2_pack(e0, e1));
    pk[j >> 1] = pack_bf16(e0, e1);
  }
  tmem_st_32x8(t_p + c * 8, pk);
}

// ---- 8-wide vector access in either activation dtype ------------------------------------------
template <typename T> struct V8;
template <> struct V8<bf16> {
  static __device__ __forceinline__ void load(const bf16* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  }
  static __device__ __forceinline__ void store(bf16* p, const float (&v)[8]) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]),
                                              pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
  // value as it will be read back (bf16 rounding applied)
  static __device__ __forceinline__ float round(float x) { return __bfloat162float(__float2bfloat16(x)); }
};
template <> struct V8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
  static __device__ __forceinline__ float round(float x) { return x; }
};
__device__ __forceinline__ float to_f32(bf16 x) { return __bfloat162float(x); }
__device__ __forceinline__ float to_f32(float x) { return x; }
template <typename T> __device__ __forceinline__ T from_f32(float x);
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float x) { return __float2bfloat16(x); }
template <> __device__ __forceinline__ float from_f32<float>(float x) { return x; }

#endif  // __CUDACC__

}  // namespace nrv
