// Micro-benchmark: FP32 FMA issue rate on sm_100a, scalar FFMA (3 register operands) vs packed fma.rn.f32x2 (FFMA2).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void fma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
template <int MODE>
__global__ void k(float* out, int iters, float s) {
  float a[16];
  unsigned long long p[8];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  for (int i = 0; i < 8; ++i) p[i] = ((unsigned long long)__float_as_uint(a[2 * i]) << 32) | __float_as_uint(a[2 * i + 1]);
  float x = s, y = s * 0.5f;
  unsigned long long xx = ((unsigned long long)__float_as_uint(x) << 32) | __float_as_uint(y), yy = xx ^ 0x1000;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, y);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) fma2(p[i], xx, yy);
    }
  }
  float r = 0;
  for (int i = 0; i < 16; ++i) r += a[i];
  for (int i = 0; i < 8; ++i) r += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  float* out; cudaMalloc(&out, sms * 8 * 256 * 4);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int warm = 0; warm < 2; ++warm) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<sms * 8, 256>>>(out, iters, 1.0001f); else k<1><<<sms * 8, 256>>>(out, iters, 1.0001f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fmas = (double)sms * 8 * 256 * iters * 16;
      if (warm) printf("%s: %.3f ms, %.1f TFLOP/s, %.1f FMA/clk/SM at %d MHz nominal\n", mode ? "FFMA2 (f32x2)" : "FFMA scalar ", ms,
                       2 * fmas / ms / 1e9, fmas / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
    }
  }
  return 0;
}
