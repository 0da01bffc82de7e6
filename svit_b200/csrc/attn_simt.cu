// Pooled attention with fused decomposed relative-position bias and residual pooling -- CUDA-core
// fp32 path (reference: slowfast/models/attention.py:429-459, :84-137, :140-183).
// Flash-style: the [Nq, Nk] score matrix is never materialised.  This is the fp32 parity mode
// (<=1e-4 per block); the bf16 production path is the tcgen05 kernel in attn_tc.cu.
//
// CTA = 32 query rows of one (batch, head); 8 warps x 4 rows; keys streamed in tiles of 64 through
// shared memory.  Bias: E[r][c] = q_r . R_c for the kh + kw + kt table rows selected by the query's
// (t, i, j) is computed once per CTA; key n = 1 + (t'*kh + i')*kw + j' then adds
// E_h[i'] + E_w[j'] + E_t[t'] -- only for patch rows x patch columns.
#include "common.cuh"
#include "../../include/svit_b200.h"

#define AQ 32
#define AK 64
#define RPW 4
#define MAXE 64
#define D SVIT_HEAD_DIM

struct AttnSmem {
  float q[AQ][D];
  float k[AK][D + 1];
  float v[AK][D];
  float e[AQ][MAXE];
  float p[AQ][AK];
};

template <typename T>
__device__ __forceinline__ void compute_bias_terms(const svit_attn_args& a, const float (*qs)[D], float (*es)[MAXE],
                                                   int64_t r0, int64_t Lq) {
  const int ne = a.kh + a.kw + a.kt;
  const T* Rh = (const T*)a.rel_h;
  const T* Rw = (const T*)a.rel_w;
  const T* Rt = (const T*)a.rel_t;
  for (int idx = threadIdx.x; idx < AQ * ne; idx += blockDim.x) {
    int r = idx / ne, c = idx % ne;
    int64_t row = r0 + r;
    float acc = 0.f;
    if (row >= 1 && row <= Lq) {
      int64_t p = row - 1;
      int j = (int)(p % a.qw), i = (int)((p / a.qw) % a.qh), t = (int)(p / ((int64_t)a.qw * a.qh));
      const T* R;
      if (c < a.kh) R = Rh + ((int64_t)i * a.kh + c) * D;
      else if (c < a.kh + a.kw) R = Rw + ((int64_t)j * a.kw + (c - a.kh)) * D;
      else R = Rt + ((int64_t)t * a.kt + (c - a.kh - a.kw)) * D;
#pragma unroll 8
      for (int d = 0; d < D; ++d) acc += qs[r][d] * to_f(R[d]);
    }
    es[r][c] = acc;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) attn_fwd_simt_kernel(svit_attn_args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  AttnSmem& s = *reinterpret_cast<AttnSmem*>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t Lq = (int64_t)a.qt * a.qh * a.qw, Lk = (int64_t)a.kt * a.kh * a.kw;
  const int64_t Nq = 1 + Lq + a.O, Nk = 1 + Lk + a.O;
  const int bh = blockIdx.y;
  const int b = bh / a.h, head = bh % a.h;
  const int64_t r0 = (int64_t)blockIdx.x * AQ;
  const T* q = (const T*)a.q + (int64_t)bh * Nq * D;
  const T* k = (const T*)a.k + (int64_t)bh * Nk * D;
  const T* v = (const T*)a.v + (int64_t)bh * Nk * D;

  for (int idx = threadIdx.x; idx < AQ * D; idx += blockDim.x) {
    int r = idx / D, d = idx % D;
    s.q[r][d] = (r0 + r < Nq) ? to_f(q[(r0 + r) * D + d]) : 0.f;
  }
  __syncthreads();
  compute_bias_terms<T>(a, s.q, s.e, r0, Lq);

  float m[RPW], l[RPW], o[RPW][3];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    m[r] = -INFINITY;
    l[r] = 0.f;
    o[r][0] = o[r][1] = o[r][2] = 0.f;
  }
  const int kh = a.kh, kw = a.kw;

  for (int64_t n0 = 0; n0 < Nk; n0 += AK) {
    __syncthreads();  // previous tile fully consumed (also orders the E writes before first use)
    for (int idx = threadIdx.x; idx < AK * D; idx += blockDim.x) {
      int n = idx / D, d = idx % D;
      bool ok = n0 + n < Nk;
      s.k[n][d] = ok ? to_f(k[(n0 + n) * D + d]) : 0.f;
      s.v[n][d] = ok ? to_f(v[(n0 + n) * D + d]) : 0.f;
    }
    __syncthreads();
    // scores for this warp's 4 rows x this lane's 2 keys
    float sc[RPW][2];
#pragma unroll
    for (int r = 0; r < RPW; ++r) sc[r][0] = sc[r][1] = 0.f;
#pragma unroll 4
    for (int d = 0; d < D; ++d) {
      float k0 = s.k[lane][d], k1 = s.k[lane + 32][d];
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        float qv = s.q[warp * RPW + r][d];
        sc[r][0] = fmaf(qv, k0, sc[r][0]);
        sc[r][1] = fmaf(qv, k1, sc[r][1]);
      }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int rr = warp * RPW + r;
      const int64_t row = r0 + rr;
      const bool qpatch = row >= 1 && row <= Lq;
      float pv[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        int64_t n = n0 + lane + 32 * u;
        float x = sc[r][u] * a.scale;
        if (qpatch && n >= 1 && n <= Lk) {
          int64_t p = n - 1;
          int jj = (int)(p % kw), ii = (int)((p / kw) % kh), tt = (int)(p / ((int64_t)kw * kh));
          x += s.e[rr][ii] + s.e[rr][kh + jj] + s.e[rr][kh + kw + tt];
        }
        pv[u] = n < Nk ? x : -INFINITY;
      }
      float mx = warp_max(fmaxf(pv[0], pv[1]));
      float mnew = fmaxf(m[r], mx);
      float corr = expf(m[r] - mnew);  // m = -inf on the first tile -> 0
      float p0 = expf(pv[0] - mnew), p1 = expf(pv[1] - mnew);
      l[r] = l[r] * corr + warp_sum(p0 + p1);
      m[r] = mnew;
      o[r][0] *= corr; o[r][1] *= corr; o[r][2] *= corr;
      s.p[rr][lane] = p0;
      s.p[rr][lane + 32] = p1;
    }
    __syncwarp();
#pragma unroll 4
    for (int n = 0; n < AK; ++n) {
      float v0 = s.v[n][lane], v1 = s.v[n][lane + 32], v2 = s.v[n][lane + 64];
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        float pp = s.p[warp * RPW + r][n];
        o[r][0] = fmaf(pp, v0, o[r][0]);
        o[r][1] = fmaf(pp, v1, o[r][1]);
        o[r][2] = fmaf(pp, v2, o[r][2]);
      }
    }
  }
  T* out = (T*)a.out;
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int rr = warp * RPW + r;
    const int64_t row = r0 + rr;
    if (row >= Nq) continue;
    float inv = 1.f / l[r];
    T* op = out + (((int64_t)b * Nq + row) * a.h + head) * D;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float val = o[r][j] * inv;
      if (row >= 1) val += s.q[rr][lane + 32 * j];  // residual pooling (attention.py:455-459)
      op[lane + 32 * j] = from_f<T>(val);
    }
    if (a.lse && lane == 0) a.lse[(int64_t)bh * Nq + row] = m[r] + logf(l[r]);
  }
}

int svit_attn_fwd_tc(const svit_attn_args* a, cudaStream_t st);  // attn_tc.cu
int svit_attn_tc_supported(const svit_attn_args* a);
int svit_attn_fwd_tc3(const svit_attn_args* a, cudaStream_t st);  // attn_tc3.cu (bias inside the score MMA)
int svit_attn_tc3_supported(const svit_attn_args* a);

static int attn_check(const svit_attn_args* a) {
  if (!a || !a->q || !a->k || !a->v || !a->out) return SVIT_EINVAL;  // the gathered tables: CUDA-core kernels only
  if (a->B < 0 || a->h < 1 || a->O < 1 || a->qt < 1 || a->qh < 1 || a->qw < 1 || a->kt < 1 || a->kh < 1 || a->kw < 1)
    return SVIT_EINVAL;
  return 0;
}

extern "C" int svit_attn_fwd(const svit_attn_args* a, void* stream) {
  int rc = attn_check(a);
  if (rc) return rc;
  if (a->B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (a->impl == 2) {
    if (svit_attn_tc3_supported(a)) return svit_attn_fwd_tc3(a, st);
    if (!svit_attn_tc_supported(a)) return SVIT_ENOTSUP;
    return svit_attn_fwd_tc(a, st);
  }
  if (a->impl == 0 && svit_attn_tc3_supported(a)) return svit_attn_fwd_tc3(a, st);
  if (a->impl == 0 && svit_attn_tc_supported(a)) return svit_attn_fwd_tc(a, st);
  if (a->kh + a->kw + a->kt > MAXE) return SVIT_ENOTSUP;
  if (!a->rel_h || !a->rel_w || !a->rel_t) return SVIT_EINVAL;  // the CUDA-core kernel reads the gathered tables
  const int64_t Nq = 1 + (int64_t)a->qt * a->qh * a->qw + a->O;
  dim3 grid((unsigned)ceil_div64(Nq, AQ), (unsigned)(a->B * a->h));
  size_t smem = sizeof(AttnSmem);
  if (a->dtype == SVIT_F32) {
    SVIT_CUDA(cudaFuncSetAttribute(attn_fwd_simt_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_fwd_simt_kernel<float><<<grid, 256, smem, st>>>(*a);
  } else if (a->dtype == SVIT_BF16) {
    SVIT_CUDA(cudaFuncSetAttribute(attn_fwd_simt_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_fwd_simt_kernel<bf16><<<grid, 256, smem, st>>>(*a);
  } else {
    return SVIT_EINVAL;
  }
  SVIT_CHECK_LAUNCH();
  return 0;
}
