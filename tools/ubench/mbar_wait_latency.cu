// How long do a failed mbarrier.try_wait and a __nanosleep(32) take on sm_100a?  (the bounded waits of the role warps)
// Also: wake-up latency of a waiter after the arrival, for three wait styles.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__global__ void k(long long* out) {
  __shared__ uint64_t bar[4];
  __shared__ long long t_arrive[4];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    bool r = try_wait(&bar[0], 0);  // fails: measures the hardware time limit
    long long t1 = clock64();
    __nanosleep(32);
    long long t2 = clock64();
    __nanosleep(32);
    long long t3 = clock64();
    bool r2 = test_wait(&bar[0], 0);
    long long t4 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = t4 - t3; out[4] = r + r2;
  }
  __syncthreads();
  // wake-up latency: warp 1 waits (style s), warp 2 arrives after a delay
  for (int s = 0; s < 3; ++s) {
    if (threadIdx.x == 32) {
      long long tw;
      if (s == 0) { while (!try_wait(&bar[1 + s], 0)) {} }
      else if (s == 1) { if (!try_wait(&bar[1 + s], 0)) { while (!try_wait(&bar[1 + s], 0)) __nanosleep(32); } }
      else { while (!test_wait(&bar[1 + s], 0)) {} }
      tw = clock64();
      out[8 + s] = tw;
    }
    if (threadIdx.x == 64) {
      long long t = clock64();
      while (clock64() - t < 20000) {}
      long long ta = clock64();
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[1 + s])) : "memory");
      t_arrive[s] = ta;
    }
    __syncthreads();
    if (threadIdx.x == 0) out[8 + s] -= t_arrive[s];
    __syncthreads();
  }
}
int main() {
  long long* d; cudaMalloc(&d, 16 * 8); cudaMemset(d, 0, 16 * 8);
  for (int it = 0; it < 3; ++it) {
    k<<<1, 96>>>(d); cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("failed try_wait %lld cyc | nanosleep(32) %lld, %lld cyc | test_wait %lld cyc | wake-up after arrive: try_wait loop %lld, try_wait+nanosleep(32) %lld, test_wait spin %lld cyc\n",
           h[0], h[1], h[2], h[3], h[8], h[9], h[10]);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
