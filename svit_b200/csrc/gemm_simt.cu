// fp32-accumulate CUDA-core GEMM with the fused epilogue of the hot path.  This is the fp32 parity
// mode (TF32 off; <=1e-4 per block against the reference) and the fallback shape handler; the bf16
// production path is the tcgen05 kernel in gemm_tc.cu, which shares the argument struct.
//
//   C[M,N] = epi( op(A)[M,K] . op(B)[K,N] )       row-major storage, op = optional transpose
//   epi(v) = residual + sample_scale[row / rows_per_sample] * ( act(v + bias) * gelu'(gelu_pre) )
//
// Used for: qkv / proj / skip-proj / fc1(+GELU) / fc2(+residual) (attention.py:345,462,561; common.py:27-34),
// the patch-embed GEMM (stem_helper.py:317) and every dgrad / wgrad of those.
#include <cstdlib>
#include "common.cuh"
#include "../../include/svit_b200.h"

#define BM 64
#define BN 64
#define BK 16

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) gemm_simt_kernel(svit_gemm_args a) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const TI* __restrict__ A = (const TI*)a.A;
  const TI* __restrict__ Bm = (const TI*)a.B;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = 0; k0 < a.K; k0 += BK) {
    // ---- A tile -> As[k][m]
    if (!a.transA) {
      int m = tid >> 2, kk = (tid & 3) * 4;
      int64_t gm = m0 + m;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int64_t gk = k0 + kk + i;
        As[kk + i][m] = (gm < a.M && gk < a.K) ? to_f(A[gm * a.lda + gk]) : 0.f;
      }
    } else {
      int kk = tid >> 4, m = (tid & 15) * 4;
      int64_t gk = k0 + kk;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int64_t gm = m0 + m + i;
        As[kk][m + i] = (gm < a.M && gk < a.K) ? to_f(A[gk * a.lda + gm]) : 0.f;
      }
    }
    // ---- B tile -> Bs[k][n]
    if (a.transB) {  // stored [N, K]
      int n = tid >> 2, kk = (tid & 3) * 4;
      int64_t gn = n0 + n;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int64_t gk = k0 + kk + i;
        Bs[kk + i][n] = (gn < a.N && gk < a.K) ? to_f(Bm[gn * a.ldb + gk]) : 0.f;
      }
    } else {  // stored [K, N]
      int kk = tid >> 4, n = (tid & 15) * 4;
      int64_t gk = k0 + kk;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int64_t gn = n0 + n + i;
        Bs[kk][n + i] = (gn < a.N && gk < a.K) ? to_f(Bm[gk * a.ldb + gn]) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- fused epilogue
  TO* __restrict__ C = (TO*)a.C;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t m = m0 + ty * 4 + i;
    if (m >= a.M) continue;
    float sc = a.sample_scale ? a.sample_scale[m / a.rows_per_sample] : 1.f;
    int64_t orow = a.rows_in > 0 ? (m / a.rows_in) * a.rows_out + a.row_off + (m % a.rows_in) : m;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t n = n0 + tx * 4 + j;
      if (n >= a.N) continue;
      float v = acc[i][j];
      if (a.bias) v += a.bias[n];
      if (a.pre_out) ((TI*)a.pre_out)[m * a.ldp + n] = from_f<TI>(v);
      if (a.act == 1) v = gelu_erf(v);
      if (a.gelu_pre) v *= gelu_erf_grad(to_f(((const TI*)a.gelu_pre)[m * a.ldg + n]));
      v *= sc;
      if (a.residual) v += to_f(((const TI*)a.residual)[orow * a.ldr + n]);
      C[orow * a.ldc + n] = from_f<TO>(v);
    }
  }
}

// column sums of a [M, N] matrix (bias gradients): out[n] += sum_m x[m, n]
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, float* __restrict__ out, int64_t M, int N, int64_t ld,
                              int64_t rows_per_cta) {
  int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
  int64_t r1 = r0 + rows_per_cta < M ? r0 + rows_per_cta : M;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int64_t r = r0; r < r1; ++r) s += to_f(x[r * ld + n]);
    atomicAdd(&out[n], s);
  }
}

// bf16, N % 8 == 0: 16-byte loads, thread = 8 columns x a stripe of rows; shared-memory reduce over the row groups,
// one atomicAdd per column and CTA
__global__ void __launch_bounds__(256) colsum_bf16x8_kernel(const bf16* __restrict__ x, float* __restrict__ out, int64_t M,
                                                            int N, int64_t ld, int64_t rows_per_cta) {
  __shared__ float red[256 * 8];
  const int n8 = N >> 3;
  const int tpr = n8 < 256 ? n8 : 256;          // threads across a row
  const int rg = 256 / tpr;                      // row groups
  const int tx = threadIdx.x % tpr, ty = threadIdx.x / tpr;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t r1 = r0 + rows_per_cta < M ? r0 + rows_per_cta : M;
  for (int c8_0 = blockIdx.x * tpr; c8_0 < n8; c8_0 += gridDim.x * tpr) {  // uniform trip count: barriers inside
    const int c8 = c8_0 + tx;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (ty < rg && c8 < n8) {
#pragma unroll 8
      for (int64_t r = r0 + ty; r < r1; r += rg) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + r * ld + c8 * 8));
        const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          acc[2 * u] += __uint_as_float(wv[u] << 16);
          acc[2 * u + 1] += __uint_as_float(wv[u] & 0xffff0000u);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) red[threadIdx.x * 8 + u] = acc[u];
    __syncthreads();
    // column sums of this CTA's column block, then atomics on consecutive addresses (a warp = one 128-byte line:
    // same-line atomics serialise in L2 at ~10 ns per transaction, so one transaction per line and CTA)
    const int ncols = min(tpr, n8 - c8_0) * 8;
    for (int n = threadIdx.x; n < ncols; n += blockDim.x) {
      const int t = n >> 3, u = n & 7;
      float sum = 0.f;
      for (int g = 0; g < rg; ++g) sum += red[(g * tpr + t) * 8 + u];
      atomicAdd(&out[c8_0 * 8 + n], sum);
    }
    __syncthreads();
  }
}

int svit_gemm_tc(const svit_gemm_args* a, cudaStream_t st);  // gemm_tc.cu
int svit_gemm_tc_supported(const svit_gemm_args* a);

// Skinny fp32 linear y = x W^T + b with few outputs (the SViTHead projections, video_model_builder.py:529-538):
// CTA per input row, the row staged in shared memory, one warp per output column (coalesced W rows, shuffle reduce).
__global__ void __launch_bounds__(256) linear_skinny_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, float* __restrict__ y, int N,
                                                            int K, int64_t lda, int64_t ldb, int64_t ldc) {
  extern __shared__ float xs[];
  const int64_t m = blockIdx.x;
  for (int k = threadIdx.x; k < K; k += blockDim.x) xs[k] = x[m * lda + k];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int n = warp; n < N; n += 8) {
    const float* wr = w + (int64_t)n * ldb;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(xs[k], __ldg(wr + k), acc);
    acc = warp_sum(acc);
    if (lane == 0) y[m * ldc + n] = acc + (bias ? bias[n] : 0.f);
  }
}

static int gemm_simt(const svit_gemm_args* a, cudaStream_t st) {
  if (a->dtype == SVIT_F32 && a->out_dtype == SVIT_F32 && !a->transA && a->transB && !a->residual && !a->gelu_pre &&
      !a->pre_out && !a->sample_scale && a->act == 0 && a->rows_in == 0 && a->M * a->N <= 65536 && a->K <= 8192) {
    linear_skinny_kernel<<<(unsigned)a->M, 256, (size_t)a->K * 4, st>>>((const float*)a->A, (const float*)a->B, a->bias,
                                                                        (float*)a->C, (int)a->N, (int)a->K, a->lda, a->ldb, a->ldc);
    SVIT_CHECK_LAUNCH();
    return 0;
  }
  dim3 grid((unsigned)ceil_div64(a->N, BN), (unsigned)ceil_div64(a->M, BM));
  if (a->dtype == SVIT_F32 && a->out_dtype == SVIT_F32)
    gemm_simt_kernel<float, float><<<grid, 256, 0, st>>>(*a);
  else if (a->dtype == SVIT_BF16 && a->out_dtype == SVIT_BF16)
    gemm_simt_kernel<bf16, bf16><<<grid, 256, 0, st>>>(*a);
  else if (a->dtype == SVIT_BF16 && a->out_dtype == SVIT_F32)
    gemm_simt_kernel<bf16, float><<<grid, 256, 0, st>>>(*a);
  else
    return SVIT_EINVAL;
  SVIT_CHECK_LAUNCH();
  return 0;
}

extern "C" {

int svit_gemm(const svit_gemm_args* a, void* stream) {
  if (!a || !a->A || !a->B || !a->C) return SVIT_EINVAL;
  if (a->M < 0 || a->N < 0 || a->K < 0) return SVIT_EINVAL;
  if (a->sample_scale && a->rows_per_sample <= 0) return SVIT_EINVAL;
  if (a->M == 0 || a->N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (a->impl == 2) {
    if (!svit_gemm_tc_supported(a)) return SVIT_ENOTSUP;
    return svit_gemm_tc(a, st);
  }
  if (a->impl == 0 && svit_gemm_tc_supported(a)) return svit_gemm_tc(a, st);
  if (a->batch > 1 || (a->alpha != 0.f && a->alpha != 1.f) || a->ln_stats) return SVIT_ENOTSUP;  // tcgen05-only features
  return gemm_simt(a, st);
}

int svit_colsum(const void* x, float* out, int64_t M, int N, int64_t ld, int dtype, void* stream) {
  if (M < 0 || N < 0) return SVIT_EINVAL;
  if (M == 0 || N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SVIT_BF16 && N % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const int n8 = N / 8, tpr = n8 < 256 ? n8 : 256;
    const unsigned gx = (unsigned)((n8 + tpr - 1) / tpr);
    int64_t rows = ceil_div64(M, ceil_div64((int64_t)svit_num_sms() * 4, gx));  // ~4 CTAs per SM
    static const int min_rows = getenv("SVIT_COLSUM_MINROWS") ? atoi(getenv("SVIT_COLSUM_MINROWS")) : 64;  // A / B switch
    if (rows < min_rows) rows = min_rows;
    dim3 grid8(gx, (unsigned)ceil_div64(M, rows));
    colsum_bf16x8_kernel<<<grid8, 256, 0, st>>>((const bf16*)x, out, M, N, ld, rows);
    SVIT_CHECK_LAUNCH();
    return 0;
  }
  int64_t rows_per_cta = 512;
  dim3 grid((unsigned)((N + 127) / 128), (unsigned)ceil_div64(M, rows_per_cta));
  if (dtype == SVIT_F32)
    colsum_kernel<float><<<grid, 128, 0, st>>>((const float*)x, out, M, N, ld, rows_per_cta);
  else if (dtype == SVIT_BF16)
    colsum_kernel<bf16><<<grid, 128, 0, st>>>((const bf16*)x, out, M, N, ld, rows_per_cta);
  else
    return SVIT_EINVAL;
  SVIT_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
