// tcgen05.ld / tcgen05.st throughput per SM on sm_100a: W warps of a CTA (warp w reads the lane quarter w % 4) issue
// back-to-back 32x32b.x32 loads (128 B per lane = 4 KB per warp instruction) of their own TMEM columns; cycles per
// instruction at 4, 8 and 16 warps per SM (1 or 2 CTAs).  The softmax warps of the attention kernels read every score
// through this path: 64 KB per (128 queries x 64 keys x 2 CTAs) round.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <bool STORE>
__global__ void __launch_bounds__(512) k(long long* out, int iters) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(256u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_ptr + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) & 1) * 32;
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
  if (STORE) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(base),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    if (STORE) {
      asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(base),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
    } else {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(base) : "memory");
      if ((it & 3) == 3) {  // keep four loads in flight
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc += r[it & 31];
      }
    }
  }
  if (STORE) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  else asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = acc + r[5]; }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_ptr), "r"(256u));
}
int main() {
  long long* out;
  cudaMalloc(&out, 4096 * sizeof(long long));
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 4096;
  for (int store = 0; store < 2; ++store)
    for (int ctas_per_sm = 1; ctas_per_sm <= 2; ++ctas_per_sm)
      for (int warps = 4; warps <= 16; warps *= 2) {
        if (warps * ctas_per_sm > 32) continue;
        const int grid = sms * ctas_per_sm;
        // 100 KB of dynamic shared memory per CTA pins the residency to the intended 1 or 2 CTAs per SM
        const size_t smem = ctas_per_sm == 1 ? 200 * 1024 : 100 * 1024;
        if (store) { cudaFuncSetAttribute(k<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<true><<<grid, warps * 32, smem>>>(out, iters); }
        else { cudaFuncSetAttribute(k<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<false><<<grid, warps * 32, smem>>>(out, iters); }
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        const double cyc = (double)h[0] / iters;  // cycles per instruction and warp
        const double bytes_per_clk = 4096.0 * warps * ctas_per_sm / cyc;
        printf("%s x32: %d CTA/SM x %2d warps: %.1f cycles per warp instruction -> %.0f B/clk/SM (%s)\n", store ? "tcgen05.st" : "tcgen05.ld",
               ctas_per_sm, warps, cyc, bytes_per_clk, cudaGetErrorString(e));
      }
  return 0;
}
