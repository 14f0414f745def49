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
#endif
