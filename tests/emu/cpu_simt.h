// Host-side SIMT emulator -- TEST INFRASTRUCTURE ONLY.
//
// Lets the CUDA kernel sources under single_speaker_tts_b200/csrc compile with g++
// (-DSSTTS_CPU_EMU) and run one thread block at a time on OS threads, so index maths,
// barriers and warp shuffles of the real kernels can be checked against the oracle in the
// CPU-only test tier (there is no GPU in the development container).  It is never linked
// into libsstts.so and nothing in the product package can reach it.
//
// Model: every CUDA thread of a block is a std::thread; __syncthreads is a block-wide
// std::barrier; each warp has its own barrier and an exchange buffer for __shfl_sync /
// __syncwarp.  Shuffles and barriers must be executed convergently (as on the device with a
// full mask).  Blocks run sequentially.
#pragma once
#include <atomic>
#include <barrier>
#include <math.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __shared__ static   /* blocks run one at a time, so function-static == block-shared */
#define SSTTS_HD inline
#define SSTTS_D inline
// the hardware faults on a vector access that is not aligned to its size; x86 does not, so the emulator checks
#define SSTTS_CHECK_ALIGNED(p, bytes)                                                                   \
  do {                                                                                                  \
    if (reinterpret_cast<uintptr_t>(p) % (bytes)) {                                                     \
      std::fprintf(stderr, "misaligned %d-byte access at %s:%d\n", (int)(bytes), __FILE__, __LINE__);   \
      std::abort();                                                                                     \
    }                                                                                                   \
  } while (0)

struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct int4 { int x, y, z, w; };
struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
// Blackwell packed FP32x2 intrinsics, emulated lane by lane (same IEEE results).
static inline float2 __fadd2_rn(float2 a, float2 b) { return float2{a.x + b.x, a.y + b.y}; }
static inline float2 __fmul2_rn(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }

namespace emu {
struct WarpCtx {
  std::unique_ptr<std::barrier<>> bar;
  uint64_t xch[32];
};
struct BlockCtx {
  std::unique_ptr<std::barrier<>> bar;
  std::vector<WarpCtx> warps;
  unsigned char* smem;
};
struct Tls {
  uint3 tid, bid;
  dim3 bdim, gdim;
  BlockCtx* blk;
  WarpCtx* warp;
};
inline Tls& tls() { static thread_local Tls t; return t; }
inline std::mutex& atomic_mutex() { static std::mutex m; return m; }

template <typename F>
void launch(dim3 grid, dim3 block, size_t smem_bytes, F body) {
  const unsigned nthreads = block.x;
  for (unsigned b = 0; b < grid.x; ++b) {
    BlockCtx blk;
    blk.bar.reset(new std::barrier<>(nthreads));
    const unsigned nwarps = (nthreads + 31) / 32;
    blk.warps.resize(nwarps);
    for (unsigned w = 0; w < nwarps; ++w) {
      unsigned cnt = std::min(32u, nthreads - w * 32);
      blk.warps[w].bar.reset(new std::barrier<>(cnt));
    }
    void* mem = nullptr;
    if (posix_memalign(&mem, 128, smem_bytes ? smem_bytes : 128) != 0) abort();
    std::memset(mem, 0xCD, smem_bytes ? smem_bytes : 128);  // poison: catch uninitialised reads
    blk.smem = static_cast<unsigned char*>(mem);
    std::vector<std::thread> ths;
    ths.reserve(nthreads);
    for (unsigned t = 0; t < nthreads; ++t) {
      ths.emplace_back([&, t]() {
        Tls& s = tls();
        s.tid = uint3{t, 0, 0};
        s.bid = uint3{b, 0, 0};
        s.bdim = block;
        s.gdim = grid;
        s.blk = &blk;
        s.warp = &blk.warps[t / 32];
        body();
      });
    }
    for (auto& th : ths) th.join();
    free(mem);
  }
}
}  // namespace emu

#define threadIdx (emu::tls().tid)
#define blockIdx (emu::tls().bid)
#define blockDim (emu::tls().bdim)
#define gridDim (emu::tls().gdim)
#define SSTTS_DYN_SMEM(name) unsigned char* name = emu::tls().blk->smem

static inline void __syncthreads() { emu::tls().blk->bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::tls().warp->bar->arrive_and_wait(); }

template <typename V>
static inline V __shfl_sync(unsigned, V v, int src) {
  static_assert(sizeof(V) <= 8, "shuffle of <= 8 byte values only");
  emu::Tls& s = emu::tls();
  uint64_t bits = 0;
  std::memcpy(&bits, &v, sizeof(V));
  s.warp->xch[s.tid.x & 31] = bits;
  s.warp->bar->arrive_and_wait();
  uint64_t got = s.warp->xch[src & 31];
  s.warp->bar->arrive_and_wait();
  V out;
  std::memcpy(&out, &got, sizeof(V));
  return out;
}
template <typename V>
static inline V __shfl_xor_sync(unsigned m, V v, int lanemask) {
  return __shfl_sync(m, v, (int)((emu::tls().tid.x & 31) ^ (unsigned)lanemask));
}
template <typename V>
static inline V __shfl_down_sync(unsigned m, V v, int delta) {
  int lane = (int)(emu::tls().tid.x & 31);
  int src = lane + delta;
  return __shfl_sync(m, v, src < 32 ? src : lane);
}

static inline void sstts_cp_async16(void* smem_dst, const void* gmem_src) {
  SSTTS_CHECK_ALIGNED(smem_dst, 16);
  SSTTS_CHECK_ALIGNED(gmem_src, 16);
  std::memcpy(smem_dst, gmem_src, 16);
}
static inline void sstts_cp_async4(void* smem_dst, const void* gmem_src) { std::memcpy(smem_dst, gmem_src, 4); }
static inline void sstts_cp_async_wait_all() {}
static inline float sstts_sqrt_approx(float x) { return sqrtf(x); }
static inline float sstts_rsqrt_approx(float x) { return 1.0f / sqrtf(x); }
static inline float sstts_log2_approx(float x) { return log2f(x); }
static inline float sstts_log2_ftz(float x) { return log2f(x); }
static inline float sstts_sin_approx(float x) { return sinf(x); }
static inline float sstts_cos_approx(float x) { return cosf(x); }
static inline void sstts_cp_async_commit() {}
static inline void sstts_prefetch_l2(const void*) {}
static inline void sstts_cp_async_wait_group1() {}
template <typename V> static inline V __ldg(const V* p) { return *p; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }

// bulk copies + mbarrier: the copy is a synchronous memcpy by the issuing thread, a completed phase bumps an
// atomic counter (sstts_mbar_phase_done, a no-op on the device where the copy engine completes the
// phase), waiters spin until the phase of the given parity is over.
typedef unsigned long long sstts_mbar_t;
static inline void sstts_mbar_init(sstts_mbar_t* bar, unsigned) { __atomic_store_n(bar, 0ULL, __ATOMIC_RELEASE); }
static inline void sstts_mbar_arrive_expect_tx(sstts_mbar_t*, unsigned) {}
static inline void sstts_bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, sstts_mbar_t*) {
  std::memcpy(smem_dst, gmem_src, bytes);
}
static inline void sstts_mbar_phase_done(sstts_mbar_t* bar) { __atomic_fetch_add(bar, 1ULL, __ATOMIC_RELEASE); }
static inline void sstts_mbar_wait(sstts_mbar_t* bar, unsigned parity) {
  while ((__atomic_load_n(bar, __ATOMIC_ACQUIRE) & 1ULL) == (unsigned long long)parity) std::this_thread::yield();
}
static inline void sstts_fence_proxy_async() {}

template <typename V> static inline V atomicAdd(V* p, V v) {
  std::lock_guard<std::mutex> g(emu::atomic_mutex());
  V old = *p; *p = old + v; return old;
}
static inline int atomicMin(int* p, int v) {
  std::lock_guard<std::mutex> g(emu::atomic_mutex());
  int old = *p; if (v < old) *p = v; return old;
}
static inline int atomicMax(int* p, int v) {
  std::lock_guard<std::mutex> g(emu::atomic_mutex());
  int old = *p; if (v > old) *p = v; return old;
}
static inline long long atomicMin(long long* p, long long v) {
  std::lock_guard<std::mutex> g(emu::atomic_mutex());
  long long old = *p; if (v < old) *p = v; return old;
}
static inline long long atomicMax(long long* p, long long v) {
  std::lock_guard<std::mutex> g(emu::atomic_mutex());
  long long old = *p; if (v > old) *p = v; return old;
}
static inline unsigned atomicMin(unsigned* p, unsigned v) {
  std::lock_guard<std::mutex> g(emu::atomic_mutex());
  unsigned old = *p; if (v < old) *p = v; return old;
}
static inline unsigned atomicMax(unsigned* p, unsigned v) {
  std::lock_guard<std::mutex> g(emu::atomic_mutex());
  unsigned old = *p; if (v > old) *p = v; return old;
}
