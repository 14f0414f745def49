// Microbenchmark: issue-slot cost of scalar FADD/FFMA vs packed FADD2/FFMA2 on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu && ./f32x2
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(128) k(float* out, int iters, float seed) {
  float2 a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = make_float2(seed + i + threadIdx.x, seed * 0.5f + i);
  const float2 c = make_float2(1.0001f, 0.9999f), d = make_float2(1e-3f, -1e-3f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { a[i].x = a[i].x + d.x; a[i].y = a[i].y + d.y; }                  // 2 FADD
      if (MODE == 1) { a[i] = __fadd2_rn(a[i], d); }                                   // 1 FADD2
      if (MODE == 2) { a[i].x = fmaf(a[i].x, c.x, d.x); a[i].y = fmaf(a[i].y, c.y, d.y); }  // 2 FFMA
      if (MODE == 3) { a[i] = __ffma2_rn(a[i], c, d); }                                // 1 FFMA2
      if (MODE == 4) {  // mixed: packed math + 2 integer-ish scalar ops competing for issue slots
        a[i] = __ffma2_rn(a[i], c, d);
        a[(i + 1) & 7].x = __int_as_float(__float_as_int(a[(i + 1) & 7].x) ^ it);
      }
      if (MODE == 5) {
        a[i].x = fmaf(a[i].x, c.x, d.x); a[i].y = fmaf(a[i].y, c.y, d.y);
        a[(i + 1) & 7].x = __int_as_float(__float_as_int(a[(i + 1) & 7].x) ^ it);
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(const char* name, int warps_per_sm) {
  float* out;
  cudaMalloc(&out, 148 * 16 * 128 * 4);
  const int iters = 20000;
  const int blocks = 148 * warps_per_sm / 4;
  k<MODE><<<blocks, 128>>>(out, 10, 1.0f);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, 128>>>(out, iters, 1.0f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  // per SMSP: warps_per_sm/4 warps, each iters*8 "pair-ops" (2 flops-lanes each)
  double pairops_per_smsp = (double)iters * 8 * warps_per_sm / 4;
  double clk = 1.965e9 * ms * 1e-3;
  printf("%-28s warps/SM %2d  %.3f ms  cycles per packed-pair-op per SMSP %.3f\n", name, warps_per_sm, ms, clk / pairops_per_smsp);
  cudaFree(out);
}

int main() {
  for (int w : {4, 16}) {
    run<0>("2x FADD", w); run<1>("1x FADD2", w); run<2>("2x FFMA", w); run<3>("1x FFMA2", w);
    run<5>("2x FFMA + LOP", w); run<4>("1x FFMA2 + LOP", w);
  }
  return 0;
}
