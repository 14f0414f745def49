// Thin portability shim for the kernel sources.
//
// Under nvcc this header only pulls in the CUDA runtime.  When SSTTS_CPU_EMU is defined (the
// host-side SIMT emulator used by tests/emu -- TEST INFRASTRUCTURE, never part of the product
// library) the emulator's header has already provided threadIdx / __syncthreads / __shfl_sync
// etc., and the kernel sources compile unchanged with g++.
#pragma once

#ifdef SSTTS_CPU_EMU
#include "cpu_simt.h"  // provided by tests/emu on the include path
#else
#include <cuda_runtime.h>
#include <stdint.h>
#define SSTTS_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define SSTTS_HD __host__ __device__ __forceinline__
#define SSTTS_D __device__ __forceinline__
// alignment assertion of vector accesses: checked by the CPU emulator build only (tests/emu/cpu_simt.h)
#define SSTTS_CHECK_ALIGNED(p, bytes) ((void)0)
// 16-byte asynchronous global -> shared copy (LDGSTS); both addresses 16-byte aligned.
__device__ __forceinline__ void sstts_cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void sstts_cp_async4(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gmem_src));
}
// pull one 128-byte line towards L2 (no register, no scoreboard entry)
__device__ __forceinline__ void sstts_prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void sstts_cp_async_commit() {
  asm volatile("cp.async.commit_group;\n" ::: "memory");
}
// wait until at most the most recent committed group is still in flight
__device__ __forceinline__ void sstts_cp_async_wait_group1() {
  asm volatile("cp.async.wait_group 1;\n" ::: "memory");
}
// SFU log2 (MUFU.LG2) and square root (MUFU.SQRT), max relative error ~2^-22
__device__ __forceinline__ float sstts_log2_approx(float x) { return __log2f(x); }
__device__ __forceinline__ float sstts_log2_ftz(float x) {       // bare MUFU.LG2 (x is never subnormal here)
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sstts_sin_approx(float x) { return __sinf(x); }   // MUFU.SIN, |x| <= pi
__device__ __forceinline__ float sstts_cos_approx(float x) { return __cosf(x); }
__device__ __forceinline__ float sstts_rsqrt_approx(float x) {   // MUFU.RSQ, denormals flushed
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sstts_sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void sstts_cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}
// ---- bulk asynchronous copies (the 1-D form of TMA) completing on a shared-memory mbarrier ----
// One thread arms the barrier with the byte count and issues the copies; the copy engine moves the data
// (no per-lane address arithmetic, no registers), everybody waits on the barrier's phase parity.
typedef unsigned long long sstts_mbar_t;
__device__ __forceinline__ void sstts_mbar_init(sstts_mbar_t* bar, unsigned arrivals) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(arrivals));
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
// the issuing thread's arrival + the number of bytes the copies of this phase will deliver
__device__ __forceinline__ void sstts_mbar_arrive_expect_tx(sstts_mbar_t* bar, unsigned bytes) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory");
}
// global -> shared bulk copy; all three of dst, src, bytes are multiples of 16
__device__ __forceinline__ void sstts_bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, sstts_mbar_t* bar) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(d), "l"(gmem_src), "r"(bytes), "r"(a) : "memory");
}
__device__ __forceinline__ void sstts_mbar_phase_done(sstts_mbar_t*) {}   // emulator hook (see cpu_simt.h)
__device__ __forceinline__ void sstts_mbar_wait(sstts_mbar_t* bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(a), "r"(parity) : "memory");
}
// orders this thread's earlier generic-proxy accesses to shared memory before later asynchronous-proxy
// (bulk copy) accesses: issued before a buffer that was read / written with ordinary instructions is
// handed back to the copy engine
__device__ __forceinline__ void sstts_fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
#endif
