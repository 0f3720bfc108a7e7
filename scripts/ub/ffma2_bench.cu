// Microbenchmark: FP32 FMA issue rate on B200 -- scalar FFMA vs packed FFMA2 (fma.rn.f32x2), the latter in
// the "scalar a  x  pair b" form the register-tiled kernels use.  Prints FMA / clk / SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float a, float2 b, float2 c) {
  unsigned long long ra, rb, rc, rd;
  float2 aa = make_float2(a, a);
  ra = *reinterpret_cast<unsigned long long*>(&aa);
  rb = *reinterpret_cast<unsigned long long*>(&b);
  rc = *reinterpret_cast<unsigned long long*>(&c);
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s0, float s1) {
  float2 acc[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) acc[q] = make_float2(threadIdx.x * 0.001f + q, q * 0.5f);
  float a = s0 + threadIdx.x * 1e-9f, b = s1;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      if (MODE == 0) { acc[q].x = fmaf(a, acc[q].x, b); acc[q].y = fmaf(a, acc[q].y, b); }
      else acc[q] = ffma2(a, acc[q], make_float2(b, b));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < 16; ++q) s += acc[q].x + acc[q].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int grid = p.multiProcessorCount * 4, iters = 20000;
  float* out; cudaMalloc(&out, grid * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<grid, 256>>>(out, iters, 0.999f, 0.001f); else k<1><<<grid, 256>>>(out, iters, 0.999f, 0.001f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
      const double fma = (double)grid * 256 * iters * 32.0;
      printf("%s rep %d: %.3f ms  %.2f TFLOP/s  %.1f FMA/clk/SM (at %d MHz nominal)\n", mode ? "FFMA2" : "FFMA ", rep, ms,
             2 * fma / ms / 1e9, fma / (ms * 1e-3) / (clk * 1e3) / p.multiProcessorCount, clk / 1000);
    }
  }
  return 0;
}
