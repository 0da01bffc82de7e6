// Backward of the pooled attention core (CUDA-core fp32 math), matching attn_simt.cu.
//   dV = P^T dO ; dP = dO V^T ; dS = P * (dP - rowsum(dP * P)) ;
//   dq = scale dS k + sum_c dE[c] R_c + dO[rows >= 1] (residual pooling) ; dk = scale dS^T q ;
//   dE_h[r][i'] = sum over patch keys with row i' of dS (likewise w, t) ; dR[a, b, :] += sum_rows dE[r][b] q_r.
// Three kernels: (1) per query tile: delta, E, dq, dE;  (2) per key tile: dk, dv;  (3) per table row: dR.
#include "common.cuh"
#include "../../include/svit_b200.h"

#define AQ 32
#define AK 64
#define RPW 4
#define MAXE 64
#define D SVIT_HEAD_DIM

struct BwdQSmem {
  float q[AQ][D];
  float dO[AQ][D];
  float k[AK][D + 1];
  float v[AK][D + 1];
  float e[AQ][MAXE];
  float de[AQ][MAXE];
  float ds[AQ][AK];
};

template <typename T>
__device__ __forceinline__ const T* rel_row(const svit_attn_args& a, int c, int i, int j, int t) {
  if (c < a.kh) return (const T*)a.rel_h + ((int64_t)i * a.kh + c) * D;
  if (c < a.kh + a.kw) return (const T*)a.rel_w + ((int64_t)j * a.kw + (c - a.kh)) * D;
  return (const T*)a.rel_t + ((int64_t)t * a.kt + (c - a.kh - a.kw)) * D;
}

template <typename T>
__global__ void __launch_bounds__(256) attn_bwd_dq_kernel(svit_attn_args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdQSmem& s = *reinterpret_cast<BwdQSmem*>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t Lq = (int64_t)a.qt * a.qh * a.qw, Lk = (int64_t)a.kt * a.kh * a.kw;
  const int64_t Nq = 1 + Lq + a.O, Nk = 1 + Lk + a.O;
  const int ne = a.kh + a.kw + a.kt;
  const int bh = blockIdx.y;
  const int b = bh / a.h, head = bh % a.h;
  const int64_t r0 = (int64_t)blockIdx.x * AQ;
  const T* q = (const T*)a.q + (int64_t)bh * Nq * D;
  const T* k = (const T*)a.k + (int64_t)bh * Nk * D;
  const T* v = (const T*)a.v + (int64_t)bh * Nk * D;
  const T* out = (const T*)a.out;
  const T* dout = (const T*)a.dout;

  for (int idx = threadIdx.x; idx < AQ * D; idx += blockDim.x) {
    int r = idx / D, d = idx % D;
    bool ok = r0 + r < Nq;
    s.q[r][d] = ok ? to_f(q[(r0 + r) * D + d]) : 0.f;
    s.dO[r][d] = ok ? to_f(dout[(((int64_t)b * Nq + r0 + r) * a.h + head) * D + d]) : 0.f;
  }
  for (int idx = threadIdx.x; idx < AQ * MAXE; idx += blockDim.x) (&s.de[0][0])[idx] = 0.f;
  __syncthreads();
  // bias terms E (same as forward) -> smem and scratch
  for (int idx = threadIdx.x; idx < AQ * ne; idx += blockDim.x) {
    int r = idx / ne, c = idx % ne;
    int64_t row = r0 + r;
    float acc = 0.f;
    if (row >= 1 && row <= Lq) {
      int64_t p = row - 1;
      int j = (int)(p % a.qw), i = (int)((p / a.qw) % a.qh), t = (int)(p / ((int64_t)a.qw * a.qh));
      const T* R = rel_row<T>(a, c, i, j, t);
#pragma unroll 8
      for (int d = 0; d < D; ++d) acc += s.q[r][d] * to_f(R[d]);
    }
    s.e[r][c] = acc;
    if (row < Nq) a.ws_e[((int64_t)bh * Nq + row) * ne + c] = acc;
  }
  // per-row lse and delta = dO . (out - residual q)
  float lse[RPW], delta[RPW], acc[RPW][3];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int rr = warp * RPW + r;
    const int64_t row = r0 + rr;
    float part = 0.f;
    if (row < Nq) {
      const T* op = out + (((int64_t)b * Nq + row) * a.h + head) * D;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        float o = to_f(op[lane + 32 * j]);
        if (row >= 1) o -= s.q[rr][lane + 32 * j];
        part += o * s.dO[rr][lane + 32 * j];
      }
    }
    delta[r] = warp_sum(part);
    lse[r] = row < Nq ? a.lse[(int64_t)bh * Nq + row] : 0.f;
    if (lane == 0 && row < Nq) a.ws_delta[(int64_t)bh * Nq + row] = delta[r];
    acc[r][0] = acc[r][1] = acc[r][2] = 0.f;
  }
  const int kh = a.kh, kw = a.kw;
  for (int64_t n0 = 0; n0 < Nk; n0 += AK) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < AK * D; idx += blockDim.x) {
      int n = idx / D, d = idx % D;
      bool ok = n0 + n < Nk;
      s.k[n][d] = ok ? to_f(k[(n0 + n) * D + d]) : 0.f;
      s.v[n][d] = ok ? to_f(v[(n0 + n) * D + d]) : 0.f;
    }
    __syncthreads();
    float sc[RPW][2], dp[RPW][2];
#pragma unroll
    for (int r = 0; r < RPW; ++r) sc[r][0] = sc[r][1] = dp[r][0] = dp[r][1] = 0.f;
#pragma unroll 2
    for (int d = 0; d < D; ++d) {
      float k0 = s.k[lane][d], k1 = s.k[lane + 32][d], v0 = s.v[lane][d], v1 = s.v[lane + 32][d];
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        float qv = s.q[warp * RPW + r][d], gv = s.dO[warp * RPW + r][d];
        sc[r][0] = fmaf(qv, k0, sc[r][0]);
        sc[r][1] = fmaf(qv, k1, sc[r][1]);
        dp[r][0] = fmaf(gv, v0, dp[r][0]);
        dp[r][1] = fmaf(gv, v1, dp[r][1]);
      }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int rr = warp * RPW + r;
      const int64_t row = r0 + rr;
      const bool qpatch = row >= 1 && row <= Lq;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        int64_t n = n0 + lane + 32 * u;
        float x = sc[r][u] * a.scale;
        bool kpatch = qpatch && n >= 1 && n <= Lk;
        int jj = 0, ii = 0, tt = 0;
        if (kpatch) {
          int64_t p = n - 1;
          jj = (int)(p % kw); ii = (int)((p / kw) % kh); tt = (int)(p / ((int64_t)kw * kh));
          x += s.e[rr][ii] + s.e[rr][kh + jj] + s.e[rr][kh + kw + tt];
        }
        float ds = 0.f;
        if (n < Nk && row < Nq) {
          float p = expf(x - lse[r]);
          ds = p * (dp[r][u] - delta[r]);
          if (kpatch) {
            atomicAdd(&s.de[rr][ii], ds);
            atomicAdd(&s.de[rr][kh + jj], ds);
            atomicAdd(&s.de[rr][kh + kw + tt], ds);
          }
        }
        s.ds[rr][lane + 32 * u] = ds;
      }
    }
    __syncwarp();
#pragma unroll 4
    for (int n = 0; n < AK; ++n) {
      float k0 = s.k[n][lane], k1 = s.k[n][lane + 32], k2 = s.k[n][lane + 64];
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        float g = s.ds[warp * RPW + r][n];
        acc[r][0] = fmaf(g, k0, acc[r][0]);
        acc[r][1] = fmaf(g, k1, acc[r][1]);
        acc[r][2] = fmaf(g, k2, acc[r][2]);
      }
    }
  }
  __syncwarp();
  T* dq = (T*)a.dq + (int64_t)bh * Nq * D;
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int rr = warp * RPW + r;
    const int64_t row = r0 + rr;
    if (row >= Nq) continue;
    float g[3] = {acc[r][0] * a.scale, acc[r][1] * a.scale, acc[r][2] * a.scale};
    if (row >= 1 && row <= Lq) {
      int64_t p = row - 1;
      int j = (int)(p % a.qw), i = (int)((p / a.qw) % a.qh), t = (int)(p / ((int64_t)a.qw * a.qh));
      for (int c = 0; c < ne; ++c) {
        float de = s.de[rr][c];
        const T* R = rel_row<T>(a, c, i, j, t);
        g[0] += de * to_f(R[lane]);
        g[1] += de * to_f(R[lane + 32]);
        g[2] += de * to_f(R[lane + 64]);
      }
    }
    if (row >= 1) {
      g[0] += s.dO[rr][lane]; g[1] += s.dO[rr][lane + 32]; g[2] += s.dO[rr][lane + 64];
    }
    dq[row * D + lane] = from_f<T>(g[0]);
    dq[row * D + lane + 32] = from_f<T>(g[1]);
    dq[row * D + lane + 64] = from_f<T>(g[2]);
    for (int c = lane; c < ne; c += 32) a.ws_de[((int64_t)bh * Nq + row) * ne + c] = s.de[rr][c];
  }
}

struct BwdKVSmem {
  float k[AK][D + 1];
  float v[AK][D + 1];
  float q[AQ][D];
  float dO[AQ][D];
  float e[AQ][MAXE];
  float p[AQ][AK];
  float ds[AQ][AK];
  float lse[AQ];
  float delta[AQ];
};

template <typename T>
__global__ void __launch_bounds__(256) attn_bwd_dkv_kernel(svit_attn_args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdKVSmem& s = *reinterpret_cast<BwdKVSmem*>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t Lq = (int64_t)a.qt * a.qh * a.qw, Lk = (int64_t)a.kt * a.kh * a.kw;
  const int64_t Nq = 1 + Lq + a.O, Nk = 1 + Lk + a.O;
  const int ne = a.kh + a.kw + a.kt;
  const int bh = blockIdx.y;
  const int b = bh / a.h, head = bh % a.h;
  const int64_t n0 = (int64_t)blockIdx.x * AK;
  const T* q = (const T*)a.q + (int64_t)bh * Nq * D;
  const T* k = (const T*)a.k + (int64_t)bh * Nk * D;
  const T* v = (const T*)a.v + (int64_t)bh * Nk * D;
  const T* dout = (const T*)a.dout;
  for (int idx = threadIdx.x; idx < AK * D; idx += blockDim.x) {
    int n = idx / D, d = idx % D;
    bool ok = n0 + n < Nk;
    s.k[n][d] = ok ? to_f(k[(n0 + n) * D + d]) : 0.f;
    s.v[n][d] = ok ? to_f(v[(n0 + n) * D + d]) : 0.f;
  }
  const int an = threadIdx.x >> 2, d0 = (threadIdx.x & 3) * 24;
  float dk[24], dv[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) dk[i] = dv[i] = 0.f;
  // key coordinates of this lane's two keys
  int kj[2], ki[2], kt_[2];
  bool kp[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    int64_t n = n0 + lane + 32 * u;
    kp[u] = n >= 1 && n <= Lk;
    int64_t p = kp[u] ? n - 1 : 0;
    kj[u] = (int)(p % a.kw); ki[u] = (int)((p / a.kw) % a.kh); kt_[u] = (int)(p / ((int64_t)a.kw * a.kh));
  }
  for (int64_t r0 = 0; r0 < Nq; r0 += AQ) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < AQ * D; idx += blockDim.x) {
      int r = idx / D, d = idx % D;
      bool ok = r0 + r < Nq;
      s.q[r][d] = ok ? to_f(q[(r0 + r) * D + d]) : 0.f;
      s.dO[r][d] = ok ? to_f(dout[(((int64_t)b * Nq + r0 + r) * a.h + head) * D + d]) : 0.f;
    }
    for (int idx = threadIdx.x; idx < AQ * ne; idx += blockDim.x) {
      int r = idx / ne, c = idx % ne;
      s.e[r][c] = r0 + r < Nq ? a.ws_e[((int64_t)bh * Nq + r0 + r) * ne + c] : 0.f;
    }
    if (threadIdx.x < AQ) {
      bool ok = r0 + threadIdx.x < Nq;
      s.lse[threadIdx.x] = ok ? a.lse[(int64_t)bh * Nq + r0 + threadIdx.x] : 0.f;
      s.delta[threadIdx.x] = ok ? a.ws_delta[(int64_t)bh * Nq + r0 + threadIdx.x] : 0.f;
    }
    __syncthreads();
    float sc[RPW][2], dp[RPW][2];
#pragma unroll
    for (int r = 0; r < RPW; ++r) sc[r][0] = sc[r][1] = dp[r][0] = dp[r][1] = 0.f;
#pragma unroll 2
    for (int d = 0; d < D; ++d) {
      float k0 = s.k[lane][d], k1 = s.k[lane + 32][d], v0 = s.v[lane][d], v1 = s.v[lane + 32][d];
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        float qv = s.q[warp * RPW + r][d], gv = s.dO[warp * RPW + r][d];
        sc[r][0] = fmaf(qv, k0, sc[r][0]);
        sc[r][1] = fmaf(qv, k1, sc[r][1]);
        dp[r][0] = fmaf(gv, v0, dp[r][0]);
        dp[r][1] = fmaf(gv, v1, dp[r][1]);
      }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int rr = warp * RPW + r;
      const int64_t row = r0 + rr;
      const bool qpatch = row >= 1 && row <= Lq;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        int64_t n = n0 + lane + 32 * u;
        float x = sc[r][u] * a.scale;
        if (qpatch && kp[u]) x += s.e[rr][ki[u]] + s.e[rr][a.kh + kj[u]] + s.e[rr][a.kh + a.kw + kt_[u]];
        float p = 0.f, ds = 0.f;
        if (n < Nk && row < Nq) {
          p = expf(x - s.lse[rr]);
          ds = p * (dp[r][u] - s.delta[rr]);
        }
        s.p[rr][lane + 32 * u] = p;
        s.ds[rr][lane + 32 * u] = ds;
      }
    }
    __syncthreads();
#pragma unroll 2
    for (int r = 0; r < AQ; ++r) {
      float pv = s.p[r][an], dsv = s.ds[r][an];
#pragma unroll
      for (int i = 0; i < 24; ++i) {
        dv[i] = fmaf(pv, s.dO[r][d0 + i], dv[i]);
        dk[i] = fmaf(dsv, s.q[r][d0 + i], dk[i]);
      }
    }
  }
  if (n0 + an < Nk) {
    T* dkp = (T*)a.dk + ((int64_t)bh * Nk + n0 + an) * D + d0;
    T* dvp = (T*)a.dv + ((int64_t)bh * Nk + n0 + an) * D + d0;
#pragma unroll
    for (int i = 0; i < 24; ++i) {
      dkp[i] = from_f<T>(dk[i] * a.scale);
      dvp[i] = from_f<T>(dv[i]);
    }
  }
}

// dR[a_idx, c, :] += sum over patch query rows whose coordinate on this axis is a_idx of dE[row][c] * q[row].
// grid = (table row = target, (b, head), row split).  Thread = one of the 96 channels x one of 4 row streams; it
// keeps all NC columns of its target in registers, so a row costs one q load, NC uniform dE loads and NC FMAs.
// The four streams are combined in shared memory and added to the fp32 table gradient with coalesced atomics.
template <typename T, int NC>
__global__ void __launch_bounds__(384) attn_bwd_drel_kernel(svit_attn_args a, int estride, int rows_per_cta) {
  __shared__ float sred[3][NC][D];
  const int d = threadIdx.x % D, g = threadIdx.x / D;  // 4 row streams
  const int64_t Lq = (int64_t)a.qt * a.qh * a.qw;
  const int64_t Nq = 1 + Lq + a.O;
  int idx = blockIdx.x, axis, kn, coff;
  float* dR;
  if (idx < a.qh) { axis = 0; kn = a.kh; coff = 0; dR = a.d_rel_h; }
  else if (idx < a.qh + a.qw) { axis = 1; idx -= a.qh; kn = a.kw; coff = a.kh; dR = a.d_rel_w; }
  else { axis = 2; idx -= a.qh + a.qw; kn = a.kt; coff = a.kh + a.kw; dR = a.d_rel_t; }
  const int n1 = axis == 2 ? a.qh : a.qt;
  const int n2 = axis == 0 ? a.qw : (axis == 1 ? a.qh : a.qw);
  const int total = n1 * n2;
  const int rb = blockIdx.z * rows_per_cta;
  if (rb >= total) return;
  const int re = min(total, rb + rows_per_cta);
  const int bh = blockIdx.y;
  const T* q = (const T*)a.q + (int64_t)bh * Nq * D + d;
  const float* de = a.ws_de + (int64_t)bh * Nq * estride + coff;
  float acc[NC];
#pragma unroll
  for (int i = 0; i < NC; ++i) acc[i] = 0.f;
#pragma unroll 2
  for (int r = rb + g; r < re; r += 4) {
    const int u = r / n2, w = r - u * n2;
    int t, i, j;
    if (axis == 0) { t = u; i = idx; j = w; }
    else if (axis == 1) { t = u; i = w; j = idx; }
    else { t = idx; i = u; j = w; }
    const int row = 1 + (t * a.qh + i) * a.qw + j;
    const float qv = to_f(q[(int64_t)row * D]);
    const float* der = de + (int64_t)row * estride;
#pragma unroll
    for (int c = 0; c < NC; ++c)
      if (c < kn) acc[c] = fmaf(der[c], qv, acc[c]);
  }
  if (g > 0) {
#pragma unroll
    for (int c = 0; c < NC; ++c) sred[g - 1][c][d] = acc[c];
  }
  __syncthreads();
  if (g == 0) {
#pragma unroll
    for (int c = 0; c < NC; ++c)
      if (c < kn) atomicAdd(&dR[((int64_t)idx * kn + c) * D + d], acc[c] + sred[0][c][d] + sred[1][c][d] + sred[2][c][d]);
  }
}

template <typename T>
static int launch_drel(const svit_attn_args* a, int estride, cudaStream_t st) {
  const int BH = a->B * a->h;
  const int kmax = max(a->kh, max(a->kw, a->kt));
  const int rmax = max(a->qt * a->qw, max(a->qt * a->qh, a->qh * a->qw));
  const int rows_per_cta = 128;
  dim3 grid(a->qh + a->qw + a->qt, BH, (rmax + rows_per_cta - 1) / rows_per_cta);
  if (BH > 65535 || grid.z > 65535) return SVIT_ENOTSUP;
  if (kmax <= 8)
    attn_bwd_drel_kernel<T, 8><<<grid, 384, 0, st>>>(*a, estride, rows_per_cta);
  else if (kmax <= 16)
    attn_bwd_drel_kernel<T, 16><<<grid, 384, 0, st>>>(*a, estride, rows_per_cta);
  else if (kmax <= 32)
    attn_bwd_drel_kernel<T, 32><<<grid, 384, 0, st>>>(*a, estride, rows_per_cta);
  else
    return SVIT_ENOTSUP;
  SVIT_CHECK_LAUNCH();
  return 0;
}

template <typename T>
static int launch_bwd(const svit_attn_args* a, cudaStream_t st) {
  const int64_t Nq = 1 + (int64_t)a->qt * a->qh * a->qw + a->O;
  const int64_t Nk = 1 + (int64_t)a->kt * a->kh * a->kw + a->O;
  const int BH = a->B * a->h;
  SVIT_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdQSmem)));
  SVIT_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdKVSmem)));
  attn_bwd_dq_kernel<T><<<dim3((unsigned)ceil_div64(Nq, AQ), BH), 256, sizeof(BwdQSmem), st>>>(*a);
  SVIT_CHECK_LAUNCH();
  attn_bwd_dkv_kernel<T><<<dim3((unsigned)ceil_div64(Nk, AK), BH), 256, sizeof(BwdKVSmem), st>>>(*a);
  SVIT_CHECK_LAUNCH();
  return launch_drel<T>(a, a->kh + a->kw + a->kt, st);
}

int svit_attn_bwd_tc(const svit_attn_args* a, cudaStream_t st);  // attn_bwd_tc.cu
int svit_attn_bwd_tc_supported(const svit_attn_args* a);

// table-gradient reduction shared with the tensor-core backward (dE rows of `estride` floats)
int svit_attn_bwd_drel(const svit_attn_args* a, int estride, cudaStream_t st) {
  if (a->dtype == SVIT_F32) return launch_drel<float>(a, estride, st);
  return launch_drel<bf16>(a, estride, st);
}

extern "C" int svit_attn_bwd(const svit_attn_args* a, void* stream) {
  if (!a || !a->q || !a->k || !a->v || !a->out || !a->lse || !a->dout || !a->dq || !a->dk || !a->dv || !a->ws_e ||
      !a->ws_de || !a->ws_delta)
    return SVIT_EINVAL;
  // gathered tables and their gradients: required unless the gradient is asked for in table-row space (d_rel_tab)
  const bool gathered = a->rel_h && a->rel_w && a->rel_t && a->d_rel_h && a->d_rel_w && a->d_rel_t;
  if (!gathered && !a->d_rel_tab) return SVIT_EINVAL;
  if (a->kh + a->kw + a->kt > MAXE) return SVIT_ENOTSUP;
  if (a->B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (a->impl == 2) return svit_attn_bwd_tc_supported(a) ? svit_attn_bwd_tc(a, st) : SVIT_ENOTSUP;
  if (a->impl == 0 && svit_attn_bwd_tc_supported(a)) return svit_attn_bwd_tc(a, st);
  if (a->d_rel_tab) return SVIT_ENOTSUP;  // the table-space gradient exists in the tensor-core path only
  if (!gathered) return SVIT_EINVAL;
  if (a->dtype == SVIT_F32) return launch_bwd<float>(a, st);
  if (a->dtype == SVIT_BF16) return launch_bwd<bf16>(a, st);
  return SVIT_EINVAL;
}
