// Backward of the pooled attention core on the tensor cores (bf16 storage, fp32 accumulate).
//
// The five contractions of the backward (attention.py:429-459 differentiated)
//     S  = q k^T            dP = dO v^T            dV = P^T dO
//     dK = scale dS^T q     dQ = scale dS k        dE = dS Sel      (rel-pos bias gradient, see below)
// run as batched problems of the tcgen05 GEMM (gemm_tc.cu: TMA-fed, TMEM accumulators, MN-major operand
// descriptors instead of transposed copies, (sample, head) two-level batch index for the head-merged dO).
// Between them three streaming kernels do the row-wise work:
//   prep      delta = dO . (out - q[rows >= 1]),  E[r, c] = q_r . R_c   (bias terms, as the forward)
//   softmax   P = exp(scale S + E[i'] + E[kh + j'] + E[kh + kw + t'] - lse),  dS = P (dP - delta)   -> bf16
//   finish    dq = scale dS k + sum_c dE[c] R_c + dO[rows >= 1] ;  dR via attn_bwd_drel (attn_simt_bwd.cu)
// The bias gradient dE[r, c] = sum over patch keys whose coordinate is c of dS[r, key] is the product of dS with the
// 0/1 key-selection matrix Sel [Nk, nep] (row = key, ones in columns i', kh + j', kh + kw + t'; zero rows for
// cls / object keys), i.e. one more GEMM with M = B h Nq.
// S and dP are kept in fp32 between the GEMMs and the softmax kernel (bf16 scores would cost 2^-9 relative error in
// the exponent argument); P and dS are the bf16 operands of the second round of GEMMs.
#include "common.cuh"
#include "../../include/svit_b200.h"

#include <cstdlib>
#define D SVIT_HEAD_DIM

int svit_gemm_tc(const svit_gemm_args* a, cudaStream_t st);  // gemm_tc.cu
int svit_gemm_tc_supported(const svit_gemm_args* a);
int svit_attn_bwd_drel(const svit_attn_args* a, int estride, cudaStream_t st);  // attn_simt_bwd.cu
int svit_attn_bwd_sdp_supported(const svit_attn_args* a);                         // attn_bwd_sdp.cu
int svit_attn_bwd_sdp(const svit_attn_args* a, cudaStream_t st);

namespace {

constexpr int PQ = 32;    // query rows per prep CTA
constexpr int MAXE = 64;  // kh + kw + kt

__device__ __forceinline__ const bf16* rel_row(const svit_attn_args& a, int c, int i, int j, int t) {
  if (c < a.kh) return (const bf16*)a.rel_h + ((int64_t)i * a.kh + c) * D;
  if (c < a.kh + a.kw) return (const bf16*)a.rel_w + ((int64_t)j * a.kw + (c - a.kh)) * D;
  return (const bf16*)a.rel_t + ((int64_t)t * a.kt + (c - a.kh - a.kw)) * D;
}

// ---- prep: delta and the bias terms E ----------------------------------------------------------------------
// 16-byte accesses throughout: a group of 16 lanes (12 active, 8 channels each) owns a row for delta; the bias terms are
// one (row, column) item per thread against the row's fp32 copy of q in shared memory (broadcast 16-byte reads).
constexpr int QP = D + 4;  // fp32 pitch of the q tile: 16-byte aligned rows

__device__ __forceinline__ void unpack8(const uint4 w, float f[8]) {
  f[0] = __uint_as_float(w.x << 16); f[1] = __uint_as_float(w.x & 0xffff0000u);
  f[2] = __uint_as_float(w.y << 16); f[3] = __uint_as_float(w.y & 0xffff0000u);
  f[4] = __uint_as_float(w.z << 16); f[5] = __uint_as_float(w.z & 0xffff0000u);
  f[6] = __uint_as_float(w.w << 16); f[7] = __uint_as_float(w.w & 0xffff0000u);
}

// etab != nullptr: the bias terms are rows of E_tab = q . T^T (fp32 [B h Nq, ldt], one tcgen05 GEMM against the
// un-gathered concatenated table, as in the forward kernel) picked through the integer index tables -- one 4-byte read
// per (row, column) instead of a 96-long dot product against a gathered table row (12 x 16-byte loads + 24 shared-memory
// reads each: the kernel was bound by them).
// e16: the bias terms are written as E / scale split into a bf16 hi + lo pair (the A-operand columns of the fused
// kernel's score product; hi + lo carries 16 mantissa bits, so the backward keeps differentiating the exact bias) into
// the ws_e scratch viewed as bf16 [B h Nq, hi (nep) | lo (nep)] -- the same bytes as its fp32 rows; else fp32 E.
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(svit_attn_args a, int nep, const float* __restrict__ etab,
                                                            int ldt, int e16) {
  __shared__ __align__(16) float sq[PQ][QP];
  const int64_t Lq = (int64_t)a.qt * a.qh * a.qw;
  const int64_t Nq = 1 + Lq + a.O;
  const int ne = a.kh + a.kw + a.kt;
  const int bh = blockIdx.y, b = bh / a.h, head = bh % a.h;
  const int64_t r0 = (int64_t)blockIdx.x * PQ;
  const bf16* q = (const bf16*)a.q + (int64_t)bh * Nq * D;
  for (int idx = threadIdx.x; idx < PQ * (D / 8); idx += blockDim.x) {
    const int r = idx / (D / 8), u = idx % (D / 8);
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (r0 + r < Nq) unpack8(__ldg(reinterpret_cast<const uint4*>(q + (r0 + r) * D) + u), f);
    *reinterpret_cast<float4*>(&sq[r][u * 8]) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(&sq[r][u * 8 + 4]) = make_float4(f[4], f[5], f[6], f[7]);
  }
  __syncthreads();
  {  // delta[row] = dO . (out - q[rows >= 1]): 16 lanes per row, two rows per group
    const int grp = threadIdx.x >> 4, l = threadIdx.x & 15;
#pragma unroll
    for (int rr = 0; rr < PQ / 16; ++rr) {
      const int r = grp * (PQ / 16) + rr;
      const int64_t row = r0 + r;
      float part = 0.f;
      if (row < Nq && l < D / 8) {
        const int64_t off = (((int64_t)b * Nq + row) * a.h + head) * D + 8 * l;
        float o[8], g[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>((const bf16*)a.out + off)), o);
        unpack8(__ldg(reinterpret_cast<const uint4*>((const bf16*)a.dout + off)), g);
        if (row >= 1) {
          const float4 qa = *reinterpret_cast<const float4*>(&sq[r][8 * l]);
          const float4 qb = *reinterpret_cast<const float4*>(&sq[r][8 * l + 4]);
          o[0] -= qa.x; o[1] -= qa.y; o[2] -= qa.z; o[3] -= qa.w;
          o[4] -= qb.x; o[5] -= qb.y; o[6] -= qb.z; o[7] -= qb.w;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) part = fmaf(o[e], g[e], part);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o, 16);
      if (l == 0 && row < Nq) a.ws_delta[(int64_t)bh * Nq + row] = part;
    }
  }
  for (int idx = threadIdx.x; idx < PQ * nep; idx += blockDim.x) {
    const int r = idx / nep, c = idx % nep;
    const int64_t row = r0 + r;
    if (row >= Nq) continue;
    float acc = 0.f;
    if (c < ne && row >= 1 && row <= Lq) {
      const int64_t p = row - 1;
      const int j = (int)(p % a.qw), i = (int)((p / a.qw) % a.qh), t = (int)(p / ((int64_t)a.qw * a.qh));
      if (etab) {
        int gidx;
        if (c < a.kh) gidx = __ldg(a.idx_h + i * a.kh + c);
        else if (c < a.kh + a.kw) gidx = a.ntab_h + __ldg(a.idx_w + j * a.kw + (c - a.kh));
        else gidx = a.ntab_h + a.ntab_w + __ldg(a.idx_t + t * a.kt + (c - a.kh - a.kw));
        acc = __ldg(etab + ((int64_t)bh * Nq + row) * ldt + gidx);
      } else {
      const uint4* R = reinterpret_cast<const uint4*>(rel_row(a, c, i, j, t));
      float acc2 = 0.f;
#pragma unroll
      for (int u = 0; u < D / 8; ++u) {
        float f[8];
        unpack8(__ldg(R + u), f);
        const float4 qa = *reinterpret_cast<const float4*>(&sq[r][u * 8]);
        const float4 qb = *reinterpret_cast<const float4*>(&sq[r][u * 8 + 4]);
        acc = fmaf(qa.x, f[0], acc); acc2 = fmaf(qa.y, f[1], acc2);
        acc = fmaf(qa.z, f[2], acc); acc2 = fmaf(qa.w, f[3], acc2);
        acc = fmaf(qb.x, f[4], acc); acc2 = fmaf(qb.y, f[5], acc2);
        acc = fmaf(qb.z, f[6], acc); acc2 = fmaf(qb.w, f[7], acc2);
      }
      acc += acc2;
      }
    }
    if (e16) {
      const float x = acc / a.scale;
      const bf16 hi = __float2bfloat16_rn(x);
      bf16* dst = reinterpret_cast<bf16*>(a.ws_e) + ((int64_t)bh * Nq + row) * 2 * nep;
      dst[c] = hi;
      dst[nep + c] = __float2bfloat16_rn(x - __bfloat162float(hi));
    } else {
      a.ws_e[((int64_t)bh * Nq + row) * nep + c] = acc;
    }
  }
}

// ---- prep, E_tab form: no shared memory, no CTA barrier, 32-bit token arithmetic -------------------------------------
// 16 lanes per (b, head, query row) over the flat row index (16 rows per CTA): delta from three 16-byte loads per lane
// and a 16-lane butterfly; the row's bias terms are 4-byte picks from its E_tab row through the integer index tables,
// written as column pairs (bf16x2 hi / lo for the fused score product, or fp32).
__global__ void __launch_bounds__(256) attn_bwd_prep_tab_kernel(svit_attn_args a, int nep, const float* __restrict__ etab,
                                                                int ldt, int e16, int64_t total_rows) {
  const int l = threadIdx.x & 15;
  const int64_t R = (int64_t)blockIdx.x * 16 + (threadIdx.x >> 4);
  const bool live = R < total_rows;
  const int64_t Rc = live ? R : 0;
  const int Lq = a.qt * a.qh * a.qw;
  const int Nq = 1 + Lq + a.O;
  const int ne = a.kh + a.kw + a.kt;
  const int bh = (int)((uint32_t)Rc / (uint32_t)Nq), row = (int)((uint32_t)Rc - (uint32_t)bh * (uint32_t)Nq);  // total_rows < 2^31
  const int b = bh / a.h, head = bh - b * a.h;
  float part = 0.f;
  if (live && l < D / 8) {
    const int64_t off = (((int64_t)b * Nq + row) * a.h + head) * D + 8 * l;
    float o[8], g[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>((const bf16*)a.out + off)), o);
    unpack8(__ldg(reinterpret_cast<const uint4*>((const bf16*)a.dout + off)), g);
    if (row >= 1) {
      float qv[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>((const bf16*)a.q + Rc * D) + l), qv);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] -= qv[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) part = fmaf(o[e], g[e], part);
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o, 16);
  if (!live) return;
  if (l == 0) a.ws_delta[Rc] = part;
  const bool patch = row >= 1 && row <= Lq;
  const int p = row - 1;
  const int j = p % a.qw, pi = p / a.qw, i = pi % a.qh, t = pi / a.qh;
  const float* erow = etab + Rc * ldt;
  for (int c2 = l; c2 < nep / 2; c2 += 16) {
    float v[2] = {0.f, 0.f};
    if (patch) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int c = 2 * c2 + u;
        if (c < ne) {
          int gidx;
          if (c < a.kh) gidx = __ldg(a.idx_h + i * a.kh + c);
          else if (c < a.kh + a.kw) gidx = a.ntab_h + __ldg(a.idx_w + j * a.kw + (c - a.kh));
          else gidx = a.ntab_h + a.ntab_w + __ldg(a.idx_t + t * a.kt + (c - a.kh - a.kw));
          v[u] = __ldg(erow + gidx);
        }
      }
    }
    if (e16) {
      // the same arithmetic as the gathered form: x = E / scale by an IEEE division, hi = rn(x), lo = rn(x - hi)
      const float x0 = v[0] / a.scale, x1 = v[1] / a.scale;
      const __nv_bfloat162 hi = __floats2bfloat162_rn(x0, x1);
      const __nv_bfloat162 lo = __floats2bfloat162_rn(x0 - __low2float(hi), x1 - __high2float(hi));
      __nv_bfloat162* dst = reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<bf16*>(a.ws_e) + Rc * 2 * nep);
      dst[c2] = hi;
      dst[nep / 2 + c2] = lo;
    } else {
      reinterpret_cast<float2*>(a.ws_e + Rc * nep)[c2] = make_float2(v[0], v[1]);
    }
  }
}

// ---- softmax recompute + dS ----------------------------------------------------------------------------------
// One warp per (b, head, query row); lanes own 4 consecutive keys per 128-key step.  Key -> E-column codes are
// built once per CTA in shared memory (slot 64 = the always-zero entry used by cls / object keys).
__global__ void __launch_bounds__(256) attn_bwd_softmax_kernel(svit_attn_args a, const float* __restrict__ S,
                                                               const float* __restrict__ dP, bf16* __restrict__ P,
                                                               bf16* __restrict__ dS, int nep, int64_t Nkp,
                                                               int64_t total_rows) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* codes = reinterpret_cast<uint32_t*>(smem_raw);  // [Nkp]
  float* es = reinterpret_cast<float*>(smem_raw + Nkp * 4);  // [8][MAXE + 4]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t Lq = (int64_t)a.qt * a.qh * a.qw, Lk = (int64_t)a.kt * a.kh * a.kw;
  const int64_t Nq = 1 + Lq + a.O, Nk = 1 + Lk + a.O;
  const int ne = a.kh + a.kw + a.kt;
  for (int64_t n = threadIdx.x; n < Nkp; n += blockDim.x) {
    uint32_t code = 64u | (64u << 8) | (64u << 16);
    if (n >= 1 && n <= Lk) {
      const int64_t p = n - 1;
      const uint32_t jj = (uint32_t)(p % a.kw), ii = (uint32_t)((p / a.kw) % a.kh), tt = (uint32_t)(p / ((int64_t)a.kw * a.kh));
      code = ii | ((a.kh + jj) << 8) | ((a.kh + a.kw + tt) << 16);
    }
    codes[n] = code;
  }
  float* e = es + warp * (MAXE + 4);
  const float kLog2e = 1.4426950408889634f;
  const float sc = a.scale * kLog2e;
  __syncthreads();
  for (int64_t R = (int64_t)blockIdx.x * 8 + warp; R < total_rows; R += (int64_t)gridDim.x * 8) {
    const int64_t row = R % Nq;
    const bool qpatch = row >= 1 && row <= Lq;
    __syncwarp();
#pragma unroll
    for (int c = lane; c < MAXE + 4; c += 32)
      e[c] = (qpatch && c < ne) ? a.ws_e[R * nep + c] * kLog2e : 0.f;
    const float nlse = -a.lse[R] * kLog2e;
    const float delta = a.ws_delta[R];
    __syncwarp();
    const float* srow = S + R * Nkp;
    const float* dprow = dP + R * Nkp;
    bf16* prow = P + R * Nkp;
    bf16* dsrow = dS + R * Nkp;
    for (int64_t n = 4 * lane; n < Nkp; n += 128) {
      const float4 s4 = __ldcs(reinterpret_cast<const float4*>(srow + n));
      const float4 d4 = __ldcs(reinterpret_cast<const float4*>(dprow + n));
      const uint4 c4 = *reinterpret_cast<const uint4*>(codes + n);
      const float sv[4] = {s4.x, s4.y, s4.z, s4.w}, dv[4] = {d4.x, d4.y, d4.z, d4.w};
      const uint32_t cv[4] = {c4.x, c4.y, c4.z, c4.w};
      float p[4], ds[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float bias = e[cv[u] & 255u] + e[(cv[u] >> 8) & 255u] + e[cv[u] >> 16];
        float pv = exp2f(fmaf(sv[u], sc, bias + nlse));
        if (n + u >= Nk) pv = 0.f;
        p[u] = pv;
        ds[u] = pv * (dv[u] - delta);
      }
      __nv_bfloat162 p01 = __floats2bfloat162_rn(p[0], p[1]), p23 = __floats2bfloat162_rn(p[2], p[3]);
      __nv_bfloat162 d01 = __floats2bfloat162_rn(ds[0], ds[1]), d23 = __floats2bfloat162_rn(ds[2], ds[3]);
      uint2 po = {*reinterpret_cast<uint32_t*>(&p01), *reinterpret_cast<uint32_t*>(&p23)};
      uint2 dso = {*reinterpret_cast<uint32_t*>(&d01), *reinterpret_cast<uint32_t*>(&d23)};
      *reinterpret_cast<uint2*>(prow + n) = po;
      *reinterpret_cast<uint2*>(dsrow + n) = dso;
    }
  }
}

// ---- G[row, g] = sum over the columns c with table row g(row, c) == g of dE[row, c]   (bf16 [B h Nq, ldg]) ----------
// The bias gradient in table-row space: with it the table term of dq is G . T and the table gradient G^T . q, two
// GEMMs instead of a 96-long FMA chain per (row, column) against gathered table rows in two CUDA-core kernels.
// ldg = table rows rounded up to 8 (pad columns are written as zeros); 16 rows per CTA, 16 lanes per row.
__global__ void __launch_bounds__(256) attn_bwd_gscatter_kernel(svit_attn_args a, int nep, bf16* __restrict__ G, int ldg,
                                                                int64_t total_rows) {
  extern __shared__ float sg_all[];  // [16][ldg]
  const int l = threadIdx.x & 15, grp = threadIdx.x >> 4;
  float* sg = sg_all + grp * ldg;
  const int64_t Lq = (int64_t)a.qt * a.qh * a.qw;
  const int64_t Nq = 1 + Lq + a.O;
  const int ne = a.kh + a.kw + a.kt;
  const int64_t R = (int64_t)blockIdx.x * 16 + grp;
  for (int i = l; i < ldg; i += 16) sg[i] = 0.f;
  __syncwarp();  // a 16-lane group only touches its own row of sg, and lies within one warp
  if (R < total_rows) {
    // 32-bit index arithmetic where the flat row index fits (ncu: four 64-bit divisions per thread made the kernel
    // issue-bound); p / (qw qh) == (p / qw) / qh for non-negative integers
    const int row = total_rows <= 0x7fffffff ? (int)((uint32_t)R % (uint32_t)Nq) : (int)(R % Nq);
    if (row >= 1 && row <= (int)Lq) {
      const int p = row - 1;
      const int jq = p % a.qw, pi = p / a.qw, iq = pi % a.qh, tq = pi / a.qh;
      for (int c = l; c < ne; c += 16) {
        int gidx;
        if (c < a.kh) gidx = __ldg(a.idx_h + iq * a.kh + c);
        else if (c < a.kh + a.kw) gidx = a.ntab_h + __ldg(a.idx_w + jq * a.kw + (c - a.kh));
        else gidx = a.ntab_h + a.ntab_w + __ldg(a.idx_t + tq * a.kt + (c - a.kh - a.kw));
        atomicAdd(&sg[gidx], __ldg(a.ws_de + R * nep + c));  // two columns may share a table row
      }
    }
  }
  __syncwarp();  // a 16-lane group only touches its own row of sg, and lies within one warp
  if (R < total_rows) {
    for (int u = l; u < ldg / 8; u += 16) {
      const float* s = &sg[8 * u];
      __nv_bfloat162 o0 = __floats2bfloat162_rn(s[0], s[1]), o1 = __floats2bfloat162_rn(s[2], s[3]);
      __nv_bfloat162 o2 = __floats2bfloat162_rn(s[4], s[5]), o3 = __floats2bfloat162_rn(s[6], s[7]);
      uint4 o;
      o.x = *reinterpret_cast<uint32_t*>(&o0); o.y = *reinterpret_cast<uint32_t*>(&o1);
      o.z = *reinterpret_cast<uint32_t*>(&o2); o.w = *reinterpret_cast<uint32_t*>(&o3);
      *(reinterpret_cast<uint4*>(G + R * ldg) + u) = o;
    }
  }
}

// ---- dq = dq_part (already scaled) + sum_c dE[c] R_c + dO[rows >= 1] -------------------------------------------
// dq_tab != nullptr: the table term arrives as a second fp32 [rows, 96] matrix (G . T), no gathers here
__global__ void __launch_bounds__(256) attn_bwd_finish_kernel(svit_attn_args a, const float* __restrict__ dq_part,
                                                              int nep, int64_t total_rows,
                                                              const float* __restrict__ dq_tab) {
  // 16 lanes per (b, head, query row), 12 of them active with 8 channels each (16-byte accesses); the row's bias
  // gradients dE[c] sit in registers of the group (lane l holds c = l, l + 16, ...) and are broadcast by shuffles
  const int l = threadIdx.x & 15;
  const int64_t Lq = (int64_t)a.qt * a.qh * a.qw;
  const int64_t Nq = 1 + Lq + a.O;
  const int ne = a.kh + a.kw + a.kt;
  const int64_t R = (int64_t)blockIdx.x * 16 + (threadIdx.x >> 4);
  const bool rok = R < total_rows;  // (all lanes stay for the shuffles)
  const int64_t Rc = rok ? R : total_rows - 1;
  const int64_t bh = Rc / Nq, row = Rc % Nq;
  const int b = (int)(bh / a.h), head = (int)(bh % a.h);
  const bool act = l < D / 8;
  float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (act) {
    const float4 ga = __ldg(reinterpret_cast<const float4*>(dq_part + Rc * D + 8 * l));
    const float4 gb = __ldg(reinterpret_cast<const float4*>(dq_part + Rc * D + 8 * l + 4));
    g[0] = ga.x; g[1] = ga.y; g[2] = ga.z; g[3] = ga.w; g[4] = gb.x; g[5] = gb.y; g[6] = gb.z; g[7] = gb.w;
  }
  const bool patch = row >= 1 && row <= Lq && !dq_tab;  // uniform within the 16-lane group
  if (dq_tab && act) {
    const float4 ta = __ldg(reinterpret_cast<const float4*>(dq_tab + Rc * D + 8 * l));
    const float4 tb = __ldg(reinterpret_cast<const float4*>(dq_tab + Rc * D + 8 * l + 4));
    g[0] += ta.x; g[1] += ta.y; g[2] += ta.z; g[3] += ta.w; g[4] += tb.x; g[5] += tb.y; g[6] += tb.z; g[7] += tb.w;
  }
  float dreg[MAXE / 16];
#pragma unroll
  for (int k = 0; k < MAXE / 16; ++k) dreg[k] = (patch && l + 16 * k < ne) ? __ldg(a.ws_de + Rc * nep + l + 16 * k) : 0.f;
  {
    const int64_t p = patch ? row - 1 : 0;
    const int jq = (int)(p % a.qw), iq = (int)((p / a.qw) % a.qh), tq = (int)(p / ((int64_t)a.qw * a.qh));
#pragma unroll
    for (int k = 0; k < MAXE / 16; ++k) {
      for (int cc = 0; cc < 16; ++cc) {
        const int c = 16 * k + cc;
        if (c >= ne) break;  // uniform
        const float w = __shfl_sync(0xffffffffu, dreg[k], cc, 16);
        if (patch && act) {
          float f[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(rel_row(a, c, iq, jq, tq)) + l), f);
#pragma unroll
          for (int e = 0; e < 8; ++e) g[e] = fmaf(w, f[e], g[e]);
        }
      }
    }
  }
  if (!rok || !act) return;
  if (row >= 1) {
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>((const bf16*)a.dout + (((int64_t)b * Nq + row) * a.h + head) * D) + l), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) g[e] += f[e];
  }
  __nv_bfloat162 o0 = __floats2bfloat162_rn(g[0], g[1]), o1 = __floats2bfloat162_rn(g[2], g[3]);
  __nv_bfloat162 o2 = __floats2bfloat162_rn(g[4], g[5]), o3 = __floats2bfloat162_rn(g[6], g[7]);
  uint4 o;
  o.x = *reinterpret_cast<uint32_t*>(&o0); o.y = *reinterpret_cast<uint32_t*>(&o1);
  o.z = *reinterpret_cast<uint32_t*>(&o2); o.w = *reinterpret_cast<uint32_t*>(&o3);
  *(reinterpret_cast<uint4*>((bf16*)a.dq + Rc * D) + l) = o;
}

// ---- finish, table-row-space form: dq = dq_part + dq_tab + dO[rows >= 1] -- a flat 8-channel-per-thread stream ----------
__global__ void __launch_bounds__(256) attn_bwd_finish_tab_kernel(svit_attn_args a, const float* __restrict__ dq_part,
                                                                  const float* __restrict__ dq_tab, uint32_t total_vecs) {
  const uint32_t v = blockIdx.x * 256u + threadIdx.x;  // (flat row, 8-channel vector); total_vecs < 2^31
  if (v >= total_vecs) return;
  const uint32_t R = v / (D / 8), u = v - R * (D / 8);
  const uint32_t Nq = 1u + (uint32_t)(a.qt * a.qh * a.qw) + (uint32_t)a.O;
  const uint32_t bh = R / Nq, row = R - bh * Nq;
  const uint32_t b = bh / (uint32_t)a.h, head = bh - b * (uint32_t)a.h;
  const float4* pa = reinterpret_cast<const float4*>(dq_part + (int64_t)R * D) + 2 * u;
  const float4* pt = reinterpret_cast<const float4*>(dq_tab + (int64_t)R * D) + 2 * u;
  const float4 ga = __ldg(pa), gb = __ldg(pa + 1), ta = __ldg(pt), tb = __ldg(pt + 1);
  float g[8] = {ga.x + ta.x, ga.y + ta.y, ga.z + ta.z, ga.w + ta.w, gb.x + tb.x, gb.y + tb.y, gb.z + tb.z, gb.w + tb.w};
  if (row >= 1) {
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>((const bf16*)a.dout + (((int64_t)b * Nq + row) * a.h + head) * D) + u), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) g[e] += f[e];
  }
  __nv_bfloat162 o0 = __floats2bfloat162_rn(g[0], g[1]), o1 = __floats2bfloat162_rn(g[2], g[3]);
  __nv_bfloat162 o2 = __floats2bfloat162_rn(g[4], g[5]), o3 = __floats2bfloat162_rn(g[6], g[7]);
  uint4 o;
  o.x = *reinterpret_cast<uint32_t*>(&o0); o.y = *reinterpret_cast<uint32_t*>(&o1);
  o.z = *reinterpret_cast<uint32_t*>(&o2); o.w = *reinterpret_cast<uint32_t*>(&o3);
  *(reinterpret_cast<uint4*>((bf16*)a.dq + (int64_t)R * D) + u) = o;
}

void gemm_defaults(svit_gemm_args& g) {
  g = svit_gemm_args{};
  g.dtype = SVIT_BF16;
  g.impl = 2;
}

int run_gemm(const svit_gemm_args& g, cudaStream_t st) {
  if (!svit_gemm_tc_supported(&g)) return SVIT_ENOTSUP;
  return svit_gemm_tc(&g, st);
}

}  // namespace

int svit_attn_bwd_tc_supported(const svit_attn_args* a) {
  if (a->dtype != SVIT_BF16) return 0;
  if (!a->ws_p || !a->ws_ds || !a->ws_dq || !a->sel_bwd) return 0;
  if ((!a->ws_s || !a->ws_dp) && !svit_attn_bwd_sdp_supported(a)) return 0;  // fp32 scratch only for the unfused path
  const int ne = a->kh + a->kw + a->kt;
  if (ne > MAXE || a->nep < ne || a->nep % 8 || a->nep > MAXE) return 0;
  const int64_t Nk = 1 + (int64_t)a->kt * a->kh * a->kw + a->O;
  if (((Nk + 7) / 8) * 8 * 4 + 8 * (MAXE + 4) * 4 > 200 * 1024) return 0;  // key-code table must fit in smem
  return 1;
}

int svit_attn_bwd_tc(const svit_attn_args* a, cudaStream_t st) {
  const int64_t Lq = (int64_t)a->qt * a->qh * a->qw, Lk = (int64_t)a->kt * a->kh * a->kw;
  const int64_t Nq = 1 + Lq + a->O, Nk = 1 + Lk + a->O;
  const int64_t Nkp = ((Nk + 7) / 8) * 8;
  const int BH = a->B * a->h, h = a->h, nep = a->nep;
  const int64_t rows = (int64_t)BH * Nq;
  int rc;

  svit_gemm_args g;
  // E_tab = q . T^T: into the (still unused) fp32 dQ scratch when the concatenated table has at most 96 rows, into the
  // caller's ws_etab (row stride = table rows rounded up to 8) beyond that
  const int ntab = a->ntab_h + a->ntab_w + a->ntab_t;
  const int ntabp = (ntab + 7) / 8 * 8;
  const float* etab = nullptr;
  int ldt = 0;
  if (a->rel_tab && a->idx_h && a->idx_w && a->idx_t && ntab >= 8 && (ntab <= D || a->ws_etab)) {
    float* dst = ntab <= D ? a->ws_dq : a->ws_etab;
    ldt = ntab <= D ? D : ntabp;
    gemm_defaults(g);
    g.A = a->q; g.lda = D;
    g.B = a->rel_tab; g.ldb = D; g.transB = 1;
    g.C = dst; g.ldc = ldt; g.out_dtype = SVIT_F32;
    g.M = rows; g.N = ntab; g.K = D; g.batch = 1;
    if (svit_gemm_tc_supported(&g)) {
      if ((rc = svit_gemm_tc(&g, st))) return rc;
      etab = dst;
    }
  }
  // table-row space: G (bf16 [rows, ntabp]) takes the place of dS, dq_tab (fp32 [rows, 96]) that of P
  const bool tab_space = a->d_rel_tab && etab && Nkp >= 2 * D && ntabp <= Nkp && ntabp <= 512;
  if (a->d_rel_tab && !tab_space) return SVIT_ENOTSUP;  // the caller would read an unwritten gradient
  if (!tab_space && !(a->rel_h && a->rel_w && a->rel_t && a->d_rel_h && a->d_rel_w && a->d_rel_t)) return SVIT_EINVAL;
  const bool fused = !(a->ws_s && a->ws_dp) && svit_attn_bwd_sdp_supported(a);
  static const bool prep_cta_form = getenv("SVIT_ATTN_BWD_PREP_CTA") != nullptr;  // A / B switch for measurements
  if (etab && rows < (int64_t)1 << 31 && !prep_cta_form)
    attn_bwd_prep_tab_kernel<<<(unsigned)ceil_div64(rows, 16), 256, 0, st>>>(*a, nep, etab, ldt, fused ? 1 : 0, rows);
  else
    attn_bwd_prep_kernel<<<dim3((unsigned)ceil_div64(Nq, PQ), BH), 256, 0, st>>>(*a, nep, etab, ldt, fused ? 1 : 0);
  SVIT_CHECK_LAUNCH();

  if (fused) {
    // S, dP, softmax and dS in one tcgen05 kernel: the fp32 matrices stay in TMEM (callers that pass the fp32 scratch
    // ask for the unfused path: key counts beyond the fused kernel's table, and the tests that compare the two)
    if ((rc = svit_attn_bwd_sdp(a, st))) return rc;
  } else {
    // S = q k^T  (fp32 [BH, Nq, Nkp])
    gemm_defaults(g);
    g.A = a->q; g.lda = D; g.strideA = Nq * D;
    g.B = a->k; g.ldb = D; g.strideB = Nk * D; g.transB = 1;
    g.C = a->ws_s; g.ldc = Nkp; g.strideC = Nq * Nkp; g.out_dtype = SVIT_F32;
    g.M = Nq; g.N = Nk; g.K = D; g.batch = BH;
    if ((rc = run_gemm(g, st))) return rc;
    // dP = dO v^T : dO is head-merged [B, Nq, h, 96] -> (sample, head) two-level batch index on A
    gemm_defaults(g);
    g.A = a->dout; g.lda = (int64_t)h * D; g.strideA = Nq * h * D; g.a_inner = h; g.strideA_inner = D;
    g.B = a->v; g.ldb = D; g.strideB = Nk * D; g.transB = 1;
    g.C = a->ws_dp; g.ldc = Nkp; g.strideC = Nq * Nkp; g.out_dtype = SVIT_F32;
    g.M = Nq; g.N = Nk; g.K = D; g.batch = BH;
    if ((rc = run_gemm(g, st))) return rc;

    {
      const size_t smem = (size_t)Nkp * 4 + 8 * (MAXE + 4) * 4;
      static SvitDevOnce configured;
      if (smem > 48 * 1024 && configured.need(smem)) {
        SVIT_CUDA(cudaFuncSetAttribute(attn_bwd_softmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.done(smem);
      }
      int64_t ctas = ceil_div64(rows, 8);
      const int64_t cap = (int64_t)svit_num_sms() * 8;
      if (ctas > cap) ctas = cap;
      attn_bwd_softmax_kernel<<<(unsigned)ctas, 256, smem, st>>>(*a, a->ws_s, a->ws_dp, (bf16*)a->ws_p, (bf16*)a->ws_ds,
                                                               nep, Nkp, rows);
      SVIT_CHECK_LAUNCH();
    }
  }

  // dV = P^T dO  (A = P stored [Nq, Nk]: MN-major A; B = dO stored [Nq, 96]: MN-major B)
  gemm_defaults(g);
  g.A = a->ws_p; g.lda = Nkp; g.strideA = Nq * Nkp; g.transA = 1;
  g.B = a->dout; g.ldb = (int64_t)h * D; g.strideB = Nq * h * D; g.b_inner = h; g.strideB_inner = D; g.transB = 0;
  g.C = a->dv; g.ldc = D; g.strideC = Nk * D; g.out_dtype = SVIT_BF16;
  g.M = Nk; g.N = D; g.K = Nq; g.batch = BH;
  if ((rc = run_gemm(g, st))) return rc;
  // dK = scale dS^T q
  gemm_defaults(g);
  g.A = a->ws_ds; g.lda = Nkp; g.strideA = Nq * Nkp; g.transA = 1;
  g.B = a->q; g.ldb = D; g.strideB = Nq * D; g.transB = 0;
  g.C = a->dk; g.ldc = D; g.strideC = Nk * D; g.out_dtype = SVIT_BF16;
  g.M = Nk; g.N = D; g.K = Nq; g.batch = BH; g.alpha = a->scale;
  if ((rc = run_gemm(g, st))) return rc;
  // dQ part = scale dS k  (fp32; the table term and the residual are added by the finish kernel)
  gemm_defaults(g);
  g.A = a->ws_ds; g.lda = Nkp; g.strideA = Nq * Nkp;
  g.B = a->k; g.ldb = D; g.strideB = Nk * D; g.transB = 0;
  g.C = a->ws_dq; g.ldc = D; g.strideC = Nq * D; g.out_dtype = SVIT_F32;
  g.M = Nq; g.N = D; g.K = Nk; g.batch = BH; g.alpha = a->scale;
  if ((rc = run_gemm(g, st))) return rc;
  // dE = dS Sel  (one problem over all B h Nq rows)
  gemm_defaults(g);
  g.A = a->ws_ds; g.lda = Nkp;
  g.B = a->sel_bwd; g.ldb = nep; g.transB = 0;
  g.C = a->ws_de; g.ldc = nep; g.out_dtype = SVIT_F32;
  g.M = rows; g.N = nep; g.K = Nk; g.batch = 1;
  if ((rc = run_gemm(g, st))) return rc;

  if (tab_space) {
    // both scratch matrices have been consumed by the GEMMs above; a row of dS holds ntabp bf16 values and a row of P
    // 96 fp32 values (checked above)
    bf16* G = (bf16*)a->ws_ds;
    float* dq_tab = (float*)a->ws_p;
    attn_bwd_gscatter_kernel<<<(unsigned)ceil_div64(rows, 16), 256, (size_t)16 * ntabp * 4, st>>>(*a, nep, G, ntabp, rows);
    SVIT_CHECK_LAUNCH();
    gemm_defaults(g);
    g.A = G; g.lda = ntabp;
    g.B = a->rel_tab; g.ldb = D; g.transB = 0;  // T stored [K = table rows, N = 96]
    g.C = dq_tab; g.ldc = D; g.out_dtype = SVIT_F32;
    g.M = rows; g.N = D; g.K = ntab; g.batch = 1;
    if ((rc = run_gemm(g, st))) return rc;
    if (rows * (D / 8) < (int64_t)1 << 31)
      attn_bwd_finish_tab_kernel<<<(unsigned)ceil_div64(rows * (D / 8), 256), 256, 0, st>>>(*a, a->ws_dq, dq_tab,
                                                                                         (uint32_t)(rows * (D / 8)));
    else
      attn_bwd_finish_kernel<<<(unsigned)ceil_div64(rows, 16), 256, 0, st>>>(*a, a->ws_dq, nep, rows, dq_tab);
    SVIT_CHECK_LAUNCH();
    // d_rel_tab = G^T . q  (split-K over the rows; the GEMM zeroes its output)
    gemm_defaults(g);
    g.A = G; g.lda = ntabp; g.transA = 1;
    g.B = a->q; g.ldb = D; g.transB = 0;
    g.C = a->d_rel_tab; g.ldc = D; g.out_dtype = SVIT_F32;
    g.M = ntab; g.N = D; g.K = rows; g.batch = 1;
    return run_gemm(g, st);
  }
  attn_bwd_finish_kernel<<<(unsigned)ceil_div64(rows, 16), 256, 0, st>>>(*a, a->ws_dq, nep, rows, nullptr);
  SVIT_CHECK_LAUNCH();
  return svit_attn_bwd_drel(a, nep, st);
}
