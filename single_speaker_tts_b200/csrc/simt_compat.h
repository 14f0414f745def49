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
#endif
