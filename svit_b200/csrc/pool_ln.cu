// attention_pool (reference: slowfast/models/attention.py:13-65) for the conv path:
//   depthwise 3x3x3 conv, stride (1,s,s), pad 1 on the patch tokens  |  cls token passes through  |
//   object tokens scaled per channel by w_eff  |  LayerNorm(96) over every token.
// Reads the packed qkv GEMM output in place ([B, N, 3, h, 96], strides passed in elements) and writes
// [B, h, N', 96]; no NCDHW round trip, no concat.  HBM-bound: ideal traffic = one read + one write.
//
// Kernel shape (round 1): one warp per output token, lane l owns channels l, l+32, l+64; the 27 taps
// are fetched through L1/L2.  The smem-tiled variant replaces this in pool_ln_tiled.cu.
#include "common.cuh"

#define PD 96
#define TAPS 27

struct PoolGeom {
  int B, h, T, H, W, Ho, Wo, O, s;
  int64_t in_bs, in_ts, in_hs;  // element strides of the input: batch, token, head
};

__device__ __forceinline__ void load_weights_to_smem(const float* __restrict__ w, const float* __restrict__ frac,
                                                     float* sw /*[27][96]*/, float* sweff /*[96]*/) {
  for (int i = threadIdx.x; i < PD * TAPS; i += blockDim.x) {
    int c = i / TAPS, t = i % TAPS;
    sw[t * PD + c] = w[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < PD; c += blockDim.x) {
    float a = 0.f;
    for (int t = 0; t < TAPS; ++t) a += sw[t * PD + c] * frac[t];
    sweff[c] = a;
  }
  __syncthreads();
}

// pre-LN value of output token `tok` (0 = cls, 1..Lo = patch, > Lo = object) for channels lane+32j
template <typename T>
__device__ __forceinline__ void pool_token(const T* __restrict__ zin, const PoolGeom& g, const float* sw,
                                           const float* sweff, int64_t tok, int lane, float v[3]) {
  const int64_t Lo = (int64_t)g.T * g.Ho * g.Wo, L = (int64_t)g.T * g.H * g.W;
  if (tok == 0) {
#pragma unroll
    for (int j = 0; j < 3; ++j) v[j] = to_f(zin[lane + 32 * j]);
  } else if (tok > Lo) {
    const T* p = zin + (tok - Lo + L) * g.in_ts;
#pragma unroll
    for (int j = 0; j < 3; ++j) v[j] = to_f(p[lane + 32 * j]) * sweff[lane + 32 * j];
  } else {
    int64_t p = tok - 1;
    int wo = (int)(p % g.Wo), ho = (int)((p / g.Wo) % g.Ho), to = (int)(p / ((int64_t)g.Wo * g.Ho));
    v[0] = v[1] = v[2] = 0.f;
#pragma unroll
    for (int kt = 0; kt < 3; ++kt) {
      int t = to - 1 + kt;
      if (t < 0 || t >= g.T) continue;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        int hh = ho * g.s - 1 + kh;
        if (hh < 0 || hh >= g.H) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          int ww = wo * g.s - 1 + kw;
          if (ww < 0 || ww >= g.W) continue;
          const T* q = zin + (1 + ((int64_t)t * g.H + hh) * g.W + ww) * g.in_ts;
          const float* wr = sw + (kt * 9 + kh * 3 + kw) * PD;
#pragma unroll
          for (int j = 0; j < 3; ++j) v[j] += to_f(q[lane + 32 * j]) * wr[lane + 32 * j];
        }
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) pool_ln_fwd_kernel(const T* __restrict__ in, PoolGeom g,
                                                          const float* __restrict__ w, const float* __restrict__ frac,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          T* __restrict__ out, float eps) {
  __shared__ float sw[TAPS * PD];
  __shared__ float sweff[PD];
  load_weights_to_smem(w, frac, sw, sweff);
  const int lane = threadIdx.x & 31;
  const int64_t Nout = 1 + (int64_t)g.T * g.Ho * g.Wo + g.O;
  const int64_t total = (int64_t)g.B * g.h * Nout;
  float gm[3], bt[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    gm[j] = gamma[lane + 32 * j];
    bt[j] = beta[lane + 32 * j];
  }
  const int wpb = blockDim.x >> 5;
  for (int64_t i = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); i < total; i += (int64_t)gridDim.x * wpb) {
    int64_t tok = i % Nout;
    int head = (int)((i / Nout) % g.h);
    int b = (int)(i / (Nout * g.h));
    const T* zin = in + b * g.in_bs + head * g.in_hs;
    float v[3];
    pool_token(zin, g, sw, sweff, tok, lane, v);
    float mean = warp_sum(v[0] + v[1] + v[2]) * (1.f / PD);
    float d0 = v[0] - mean, d1 = v[1] - mean, d2 = v[2] - mean;
    float rstd = rsqrtf(warp_sum(d0 * d0 + d1 * d1 + d2 * d2) * (1.f / PD) + eps);
    T* o = out + i * PD;
    o[lane] = from_f<T>(d0 * rstd * gm[0] + bt[0]);
    o[lane + 32] = from_f<T>(d1 * rstd * gm[1] + bt[1]);
    o[lane + 64] = from_f<T>(d2 * rstd * gm[2] + bt[2]);
  }
}

// Backward, output-centric half: recompute pre-LN value, LayerNorm backward -> dpre (written to `dpre`),
// accumulate dgamma, dbeta, and the conv weight gradient dw[c][tap] += dpre[c] * z[tap][c]
// (+ frac[tap] * sum_obj dpre[c] z_obj[c] for the object-token scale path).
template <typename T>
__global__ void __launch_bounds__(256) pool_ln_bwd_out_kernel(const T* __restrict__ in, PoolGeom g,
                                                              const float* __restrict__ w, const float* __restrict__ frac,
                                                              const float* __restrict__ gamma, const T* __restrict__ dout,
                                                              T* __restrict__ dpre, float* __restrict__ dw,
                                                              float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                              float eps) {
  __shared__ float sw[TAPS * PD];
  __shared__ float sweff[PD];
  __shared__ float sacc[(TAPS + 3) * PD];  // dw | dweff | dgamma | dbeta
  load_weights_to_smem(w, frac, sw, sweff);
  for (int i = threadIdx.x; i < (TAPS + 3) * PD; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t Lo = (int64_t)g.T * g.Ho * g.Wo, L = (int64_t)g.T * g.H * g.W;
  const int64_t Nout = 1 + Lo + g.O;
  const int64_t total = (int64_t)g.B * g.h * Nout;
  float gm[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) gm[j] = gamma[lane + 32 * j];
  float aw[TAPS][3];
  float aweff[3] = {0.f, 0.f, 0.f}, ag[3] = {0.f, 0.f, 0.f}, ab[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int t = 0; t < TAPS; ++t) aw[t][0] = aw[t][1] = aw[t][2] = 0.f;
  const int wpb = blockDim.x >> 5;
  for (int64_t i = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); i < total; i += (int64_t)gridDim.x * wpb) {
    int64_t tok = i % Nout;
    int head = (int)((i / Nout) % g.h);
    int b = (int)(i / (Nout * g.h));
    const T* zin = in + b * g.in_bs + head * g.in_hs;
    float v[3];
    pool_token(zin, g, sw, sweff, tok, lane, v);
    float mean = warp_sum(v[0] + v[1] + v[2]) * (1.f / PD);
    float xh[3], gy[3], dp[3];
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      xh[j] = v[j] - mean;
      q += xh[j] * xh[j];
    }
    float rstd = rsqrtf(warp_sum(q) * (1.f / PD) + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      xh[j] *= rstd;
      float d = to_f(dout[i * PD + lane + 32 * j]);
      gy[j] = d * gm[j];
      s1 += gy[j];
      s2 += gy[j] * xh[j];
      ag[j] += d * xh[j];
      ab[j] += d;
    }
    s1 = warp_sum(s1) * (1.f / PD);
    s2 = warp_sum(s2) * (1.f / PD);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      dp[j] = rstd * (gy[j] - s1 - xh[j] * s2);
      dpre[i * PD + lane + 32 * j] = from_f<T>(dp[j]);
    }
    if (tok > Lo) {
      const T* p = zin + (tok - Lo + L) * g.in_ts;
#pragma unroll
      for (int j = 0; j < 3; ++j) aweff[j] += dp[j] * to_f(p[lane + 32 * j]);
    } else if (tok > 0) {
      int64_t p = tok - 1;
      int wo = (int)(p % g.Wo), ho = (int)((p / g.Wo) % g.Ho), to = (int)(p / ((int64_t)g.Wo * g.Ho));
#pragma unroll
      for (int kt = 0; kt < 3; ++kt) {
        int t = to - 1 + kt;
        if (t < 0 || t >= g.T) continue;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          int hh = ho * g.s - 1 + kh;
          if (hh < 0 || hh >= g.H) continue;
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            int ww = wo * g.s - 1 + kw;
            if (ww < 0 || ww >= g.W) continue;
            const T* qz = zin + (1 + ((int64_t)t * g.H + hh) * g.W + ww) * g.in_ts;
#pragma unroll
            for (int j = 0; j < 3; ++j) aw[kt * 9 + kh * 3 + kw][j] += dp[j] * to_f(qz[lane + 32 * j]);
          }
        }
      }
    }
  }
  // CTA reduction in shared memory, then one global atomic per entry per CTA
#pragma unroll
  for (int t = 0; t < TAPS; ++t)
#pragma unroll
    for (int j = 0; j < 3; ++j) atomicAdd(&sacc[t * PD + lane + 32 * j], aw[t][j]);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    atomicAdd(&sacc[TAPS * PD + lane + 32 * j], aweff[j]);
    atomicAdd(&sacc[(TAPS + 1) * PD + lane + 32 * j], ag[j]);
    atomicAdd(&sacc[(TAPS + 2) * PD + lane + 32 * j], ab[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < TAPS * PD; i += blockDim.x) {
    int t = i / PD, c = i % PD;
    atomicAdd(&dw[c * TAPS + t], sacc[i] + frac[t] * sacc[TAPS * PD + c]);
  }
  for (int c = threadIdx.x; c < PD; c += blockDim.x) {
    atomicAdd(&dgamma[c], sacc[(TAPS + 1) * PD + c]);
    atomicAdd(&dbeta[c], sacc[(TAPS + 2) * PD + c]);
  }
}

// Backward, input-centric half: dz[token] = transposed depthwise conv of dpre (patch), dpre (cls),
// dpre * w_eff (object tokens).  Written with the input strides (i.e. straight into the dqkv buffer).
template <typename T>
__global__ void __launch_bounds__(256) pool_ln_bwd_in_kernel(const T* __restrict__ dpre, PoolGeom g,
                                                             const float* __restrict__ w, const float* __restrict__ frac,
                                                             T* __restrict__ dz) {
  __shared__ float sw[TAPS * PD];
  __shared__ float sweff[PD];
  load_weights_to_smem(w, frac, sw, sweff);
  const int lane = threadIdx.x & 31;
  const int64_t Lo = (int64_t)g.T * g.Ho * g.Wo, L = (int64_t)g.T * g.H * g.W;
  const int64_t Nout = 1 + Lo + g.O, Nin = 1 + L + g.O;
  const int64_t total = (int64_t)g.B * g.h * Nin;
  const int wpb = blockDim.x >> 5;
  for (int64_t i = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); i < total; i += (int64_t)gridDim.x * wpb) {
    int64_t tok = i % Nin;
    int head = (int)((i / Nin) % g.h);
    int b = (int)(i / (Nin * g.h));
    const T* dp = dpre + ((int64_t)b * g.h + head) * Nout * PD;
    float v[3] = {0.f, 0.f, 0.f};
    if (tok == 0) {
#pragma unroll
      for (int j = 0; j < 3; ++j) v[j] = to_f(dp[lane + 32 * j]);
    } else if (tok > L) {
      const T* p = dp + (tok - L + Lo) * PD;
#pragma unroll
      for (int j = 0; j < 3; ++j) v[j] = to_f(p[lane + 32 * j]) * sweff[lane + 32 * j];
    } else {
      int64_t p = tok - 1;
      int ww = (int)(p % g.W), hh = (int)((p / g.W) % g.H), t = (int)(p / ((int64_t)g.W * g.H));
#pragma unroll
      for (int kt = 0; kt < 3; ++kt) {
        int to = t + 1 - kt;
        if (to < 0 || to >= g.T) continue;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          int num = hh + 1 - kh;
          if (num < 0 || num % g.s != 0) continue;
          int ho = num / g.s;
          if (ho >= g.Ho) continue;
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            int numw = ww + 1 - kw;
            if (numw < 0 || numw % g.s != 0) continue;
            int wo = numw / g.s;
            if (wo >= g.Wo) continue;
            const T* q = dp + (1 + ((int64_t)to * g.Ho + ho) * g.Wo + wo) * PD;
            const float* wr = sw + (kt * 9 + kh * 3 + kw) * PD;
#pragma unroll
            for (int j = 0; j < 3; ++j) v[j] += to_f(q[lane + 32 * j]) * wr[lane + 32 * j];
          }
        }
      }
    }
    T* o = dz + b * g.in_bs + head * g.in_hs + tok * g.in_ts;
#pragma unroll
    for (int j = 0; j < 3; ++j) o[lane + 32 * j] = from_f<T>(v[j]);
  }
}

int svit_pool_ln_fwd_bf16(const void* in, int64_t in_bs, int64_t in_ts, int64_t in_hs, const float* conv_w,
                          const float* tap_frac, const float* gamma, const float* beta, void* out, int B, int h, int T,
                          int H, int W, int O, int s, float eps, cudaStream_t st, void* pre);  // pool_ln_tiled.cu

int svit_pool_ln_bwd_bf16_supported(const void* in, int64_t in_bs, int64_t in_ts, int64_t in_hs, const void* dout,
                                    const void* dpre, const void* dz);  // pool_ln_bwd_bf16.cu
int svit_pool_ln_bwd_bf16(const void* in, int64_t in_bs, int64_t in_ts, int64_t in_hs, const float* conv_w,
                          const float* tap_frac, const float* gamma, const void* dout, void* dpre, void* dz, float* dw,
                          float* dgamma, float* dbeta, int B, int h, int T, int H, int W, int O, int s, float eps,
                          cudaStream_t st, const void* pre);

static int make_geom(PoolGeom& g, int B, int h, int T, int H, int W, int O, int s, int64_t in_bs, int64_t in_ts,
                     int64_t in_hs) {
  if (B < 0 || h < 1 || T < 1 || H < 1 || W < 1 || O < 1 || s < 1) return SVIT_EINVAL;
  g.B = B; g.h = h; g.T = T; g.H = H; g.W = W; g.O = O; g.s = s;
  g.Ho = (H - 1) / s + 1;  // floor((H + 2 - 3) / s) + 1
  g.Wo = (W - 1) / s + 1;
  g.in_bs = in_bs; g.in_ts = in_ts; g.in_hs = in_hs;
  return 0;
}

static inline int pool_grid(int64_t tokens) {
  int64_t g = ceil_div64(tokens, 8);
  int64_t cap = (int64_t)svit_num_sms() * 8;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

extern "C" {

int svit_pool_ln_fwd(const void* in, int64_t in_batch_stride, int64_t in_tok_stride, int64_t in_head_stride,
                     const float* conv_w, const float* tap_frac, const float* gamma, const float* beta, void* out,
                     int B, int h, int T, int H, int W, int O, int stride_hw, float eps, int dtype, void* stream) {
  PoolGeom g;
  int rc = make_geom(g, B, h, T, H, W, O, stride_hw, in_batch_stride, in_tok_stride, in_head_stride);
  if (rc) return rc;
  int64_t tokens = (int64_t)B * h * (1 + (int64_t)T * g.Ho * g.Wo + O);
  if (tokens == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SVIT_BF16 && in_batch_stride % 2 == 0 && in_tok_stride % 2 == 0 && in_head_stride % 2 == 0 &&
      (reinterpret_cast<uintptr_t>(in) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0)
    return svit_pool_ln_fwd_bf16(in, in_batch_stride, in_tok_stride, in_head_stride, conv_w, tap_frac, gamma, beta, out,
                                 B, h, T, H, W, O, stride_hw, eps, st, nullptr);
  if (dtype == SVIT_F32)
    pool_ln_fwd_kernel<float><<<pool_grid(tokens), 256, 0, st>>>((const float*)in, g, conv_w, tap_frac, gamma, beta, (float*)out, eps);
  else if (dtype == SVIT_BF16)
    pool_ln_fwd_kernel<bf16><<<pool_grid(tokens), 256, 0, st>>>((const bf16*)in, g, conv_w, tap_frac, gamma, beta, (bf16*)out, eps);
  else
    return SVIT_EINVAL;
  SVIT_CHECK_LAUNCH();
  return 0;
}

// dw [96*27], dgamma[96], dbeta[96] are accumulated into (+=); dpre is caller-provided scratch of the
// output's shape; dz is written with the same strides as `in`.
int svit_pool_ln_bwd(const void* in, int64_t in_batch_stride, int64_t in_tok_stride, int64_t in_head_stride,
                     const float* conv_w, const float* tap_frac, const float* gamma, const void* dout, void* dpre,
                     void* dz, float* dw, float* dgamma, float* dbeta, int B, int h, int T, int H, int W, int O,
                     int stride_hw, float eps, int dtype, void* stream) {
  PoolGeom g;
  int rc = make_geom(g, B, h, T, H, W, O, stride_hw, in_batch_stride, in_tok_stride, in_head_stride);
  if (rc) return rc;
  int64_t tok_out = (int64_t)B * h * (1 + (int64_t)T * g.Ho * g.Wo + O);
  int64_t tok_in = (int64_t)B * h * (1 + (int64_t)T * H * W + O);
  if (tok_out == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SVIT_BF16 && svit_pool_ln_bwd_bf16_supported(in, in_batch_stride, in_tok_stride, in_head_stride, dout, dpre, dz))
    return svit_pool_ln_bwd_bf16(in, in_batch_stride, in_tok_stride, in_head_stride, conv_w, tap_frac, gamma, dout, dpre,
                                 dz, dw, dgamma, dbeta, B, h, T, H, W, O, stride_hw, eps, st, nullptr);
  int g1 = pool_grid(tok_out);
  if (g1 > svit_num_sms() * 2) g1 = svit_num_sms() * 2;
  if (dtype == SVIT_F32) {
    pool_ln_bwd_out_kernel<float><<<g1, 256, 0, st>>>((const float*)in, g, conv_w, tap_frac, gamma, (const float*)dout, (float*)dpre, dw, dgamma, dbeta, eps);
    pool_ln_bwd_in_kernel<float><<<pool_grid(tok_in), 256, 0, st>>>((const float*)dpre, g, conv_w, tap_frac, (float*)dz);
  } else if (dtype == SVIT_BF16) {
    pool_ln_bwd_out_kernel<bf16><<<g1, 256, 0, st>>>((const bf16*)in, g, conv_w, tap_frac, gamma, (const bf16*)dout, (bf16*)dpre, dw, dgamma, dbeta, eps);
    pool_ln_bwd_in_kernel<bf16><<<pool_grid(tok_in), 256, 0, st>>>((const bf16*)dpre, g, conv_w, tap_frac, (bf16*)dz);
  } else {
    return SVIT_EINVAL;
  }
  SVIT_CHECK_LAUNCH();
  return 0;
}

// Training variants (bf16): the forward also writes the pre-LayerNorm rows (`pre`, same shape as out) and the backward
// reads them instead of recomputing the 27-tap convolution per output token (a third of the pooling backward).
int svit_pool_ln_fwd_save(const void* in, int64_t in_batch_stride, int64_t in_tok_stride, int64_t in_head_stride,
                          const float* conv_w, const float* tap_frac, const float* gamma, const float* beta, void* out,
                          void* pre, int B, int h, int T, int H, int W, int O, int stride_hw, float eps, int dtype,
                          void* stream) {
  PoolGeom g;
  int rc = make_geom(g, B, h, T, H, W, O, stride_hw, in_batch_stride, in_tok_stride, in_head_stride);
  if (rc) return rc;
  if (!pre) return SVIT_EINVAL;
  if (dtype != SVIT_BF16 || in_batch_stride % 2 || in_tok_stride % 2 || in_head_stride % 2 ||
      ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(pre)) & 15))
    return SVIT_ENOTSUP;
  if ((int64_t)B * h * (1 + (int64_t)T * g.Ho * g.Wo + O) == 0) return 0;
  return svit_pool_ln_fwd_bf16(in, in_batch_stride, in_tok_stride, in_head_stride, conv_w, tap_frac, gamma, beta, out, B, h,
                               T, H, W, O, stride_hw, eps, (cudaStream_t)stream, pre);
}

int svit_pool_ln_bwd_saved(const void* in, int64_t in_batch_stride, int64_t in_tok_stride, int64_t in_head_stride,
                           const float* conv_w, const float* tap_frac, const float* gamma, const void* dout,
                           const void* pre, void* dpre, void* dz, float* dw, float* dgamma, float* dbeta, int B, int h,
                           int T, int H, int W, int O, int stride_hw, float eps, int dtype, void* stream) {
  PoolGeom g;
  int rc = make_geom(g, B, h, T, H, W, O, stride_hw, in_batch_stride, in_tok_stride, in_head_stride);
  if (rc) return rc;
  if (!pre) return SVIT_EINVAL;
  if (dtype != SVIT_BF16 ||
      !svit_pool_ln_bwd_bf16_supported(in, in_batch_stride, in_tok_stride, in_head_stride, dout, dpre, dz) ||
      (reinterpret_cast<uintptr_t>(pre) & 7))
    return SVIT_ENOTSUP;
  if ((int64_t)B * h * (1 + (int64_t)T * g.Ho * g.Wo + O) == 0) return 0;
  return svit_pool_ln_bwd_bf16(in, in_batch_stride, in_tok_stride, in_head_stride, conv_w, tap_frac, gamma, dout, dpre, dz,
                               dw, dgamma, dbeta, B, h, T, H, W, O, stride_hw, eps, (cudaStream_t)stream, pre);
}

}  // extern "C"
