// Shared device/host helpers for the svit_b200 sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define SVIT_F32 0
#define SVIT_BF16 1

#define SVIT_HEAD_DIM 96

// Error convention of the C ABI (include/svit_b200.h): 0 ok, <0 argument error, >0 cudaError_t.
#define SVIT_EINVAL (-1)
#define SVIT_ENOTSUP (-2)

#define SVIT_CHECK_LAUNCH()                                  \
  do {                                                       \
    cudaError_t e__ = cudaGetLastError();                    \
    if (e__ != cudaSuccess) return (int)e__;                 \
  } while (0)

#define SVIT_CUDA(x)                                         \
  do {                                                       \
    cudaError_t e__ = (x);                                   \
    if (e__ != cudaSuccess) return (int)e__;                 \
  } while (0)

typedef __nv_bfloat16 bf16;

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

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

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float kInvSqrt2Pi = 0.39894228040143267794f;
  float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  return cdf + x * kInvSqrt2Pi * __expf(-0.5f * x * x);
}

static inline int svit_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
