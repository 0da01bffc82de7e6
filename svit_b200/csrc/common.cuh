// Shared device/host helpers for the svit_b200 sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

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

#define SVIT_MAX_DEVICES 64

static inline int svit_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev >= 0 && dev < SVIT_MAX_DEVICES ? dev : 0;
}

// SM count of the CURRENT device (a process may drive several GPUs).
static inline int svit_num_sms() {
  static std::atomic<int> n[SVIT_MAX_DEVICES];
  const int dev = svit_device();
  int v = n[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
    n[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

// One-time kernel configuration PER DEVICE.  cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the current
// device's context only, so a process-wide flag would leave the second GPU of a process unconfigured.
//   static SvitDevOnce once;  if (once.need(bytes)) { cudaFuncSetAttribute(...); once.done(bytes); }
// need(v): the current device has not been configured with a value >= v yet.  Threads racing on the same device may
// both configure (the call is idempotent); the flag is published only after the attribute has been set.
struct SvitDevOnce {
  std::atomic<size_t> v[SVIT_MAX_DEVICES];
  bool need(size_t want = 1) { return v[svit_device()].load(std::memory_order_acquire) < want; }
  void done(size_t want = 1) {
    std::atomic<size_t>& a = v[svit_device()];
    size_t cur = a.load(std::memory_order_relaxed);
    while (cur < want && !a.compare_exchange_weak(cur, want, std::memory_order_release)) {
    }
  }
};

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
