// Backward of attention_pool (conv pool + LayerNorm, attention.py:13-65) for bf16 activations, vectorised.
// Three kernels, all reading / writing the packed qkv layout in place:
//   pre   warp per OUTPUT token, lanes 0..23 own 4 consecutive channels (8-byte accesses): recompute the pre-LN
//         conv value, LayerNorm backward -> dpre (bf16 scratch), dgamma / dbeta partial sums
//   dw    thread = (channel pair, temporal tap plane, token parity) marching over a chunk of output tokens:
//         dw[c][tap] += dpre[tok][c] * z[tok + tap][c] in 18 registers, plus the object-token path
//         (d w_eff = sum_obj dpre * z_obj, d w[c][tap] += tap_frac[tap] * d w_eff[c]); one atomic per entry and CTA
//   in    warp per INPUT token: dz = transposed depthwise conv of dpre (patch), dpre (cls), dpre * w_eff (object)
// HBM-bound in principle (two reads + two writes of the tensors); in practice bound by L2 latency, so the kernels
// favour wide loads and many resident warps over register-resident accumulators.
#include "common.cuh"

#define PD 96
#define TAPS 27

namespace {

struct Geom {
  int B, h, T, H, W, Ho, Wo, O, s;
  int64_t in_bs, in_ts, in_hs;
  int its;  // in_ts as int: offsets inside one (b, head) slice fit 32 bits (checked at launch)
};

__device__ __forceinline__ void load_weights(const float* __restrict__ w, const float* __restrict__ frac, float* sw,
                                             float* sweff) {
  for (int i = threadIdx.x; i < PD * TAPS; i += blockDim.x) {
    const int c = i / TAPS, t = i % TAPS;
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

__device__ __forceinline__ void unpack4(const uint2 v, float f[4]) {
  f[0] = __uint_as_float(v.x << 16);
  f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16);
  f[3] = __uint_as_float(v.y & 0xffff0000u);
}
__device__ __forceinline__ uint2 pack4(const float f[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
  return make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}

__global__ void __launch_bounds__(256) pool_bwd_pre_kernel(const bf16* __restrict__ in, Geom g,
                                                           const float* __restrict__ w, const float* __restrict__ frac,
                                                           const float* __restrict__ gamma, const bf16* __restrict__ dout,
                                                           bf16* __restrict__ dpre, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta, float eps,
                                                           const bf16* __restrict__ pre) {
  __shared__ __align__(16) float sw[TAPS * PD];
  __shared__ __align__(16) float sweff[PD];
  __shared__ float sred[2 * PD];
  if (!pre) load_weights(w, frac, sw, sweff);  // the saved pre-LayerNorm rows make the convolution (and its weights) unnecessary
  for (int i = threadIdx.x; i < 2 * PD; i += blockDim.x) sred[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const bool act = lane < 24;
  const int c0 = act ? lane * 4 : 0;
  const int Lo = g.T * g.Ho * g.Wo, L = g.T * g.H * g.W;
  const int Nout = 1 + Lo + g.O;
  const float4 gm = *reinterpret_cast<const float4*>(gamma + c0);
  float ag[4] = {0.f, 0.f, 0.f, 0.f}, ab[4] = {0.f, 0.f, 0.f, 0.f};
  const int wpb = blockDim.x >> 5;
  const int bh = blockIdx.y, head = bh % g.h, b = bh / g.h;
  const bf16* zin = in + b * g.in_bs + head * g.in_hs + c0;
  // patch-token coordinates advance with the token stride (decomposed once): no division in the loop
  const int step = gridDim.x * wpb;
  const int step_w = step % g.Wo, step_h = (step / g.Wo) % g.Ho, step_t = step / (g.Wo * g.Ho);
  int wo, ho, to;
  {
    const int tok0 = blockIdx.x * wpb + (threadIdx.x >> 5);
    const unsigned p = (tok0 < 1 ? tok0 + step : tok0) - 1;
    const unsigned pr = p / (unsigned)g.Wo;
    wo = (int)(p - pr * g.Wo); to = (int)(pr / (unsigned)g.Ho); ho = (int)(pr - to * g.Ho);
  }
  for (int tok = blockIdx.x * wpb + (threadIdx.x >> 5); tok < Nout; tok += step) {
    const int64_t i = (int64_t)bh * Nout + tok;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    float dy[4];
    unpack4(act ? __ldg(reinterpret_cast<const uint2*>(dout + i * PD + c0)) : make_uint2(0, 0), dy);
    if (act && pre) {  // the forward saved the pre-LayerNorm row: no convolution to recompute
      unpack4(__ldg(reinterpret_cast<const uint2*>(pre + i * PD + c0)), v);
    } else if (act) {
      if (tok == 0) {
        unpack4(__ldg(reinterpret_cast<const uint2*>(zin)), v);
      } else if (tok > Lo) {
        float z[4];
        unpack4(__ldg(reinterpret_cast<const uint2*>(zin + (tok - Lo + L) * g.its)), z);
        const float4 we = *reinterpret_cast<const float4*>(sweff + c0);
        v[0] = z[0] * we.x; v[1] = z[1] * we.y; v[2] = z[2] * we.z; v[3] = z[3] * we.w;
      } else {
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
          const int t = to - 1 + kt;
          if (t < 0 || t >= g.T) continue;
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const int hh = ho * g.s - 1 + kh;
            if (hh < 0 || hh >= g.H) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const int ww = wo * g.s - 1 + kw;
              if (ww < 0 || ww >= g.W) continue;
              float z[4];
              unpack4(__ldg(reinterpret_cast<const uint2*>(zin + (1 + (t * g.H + hh) * g.W + ww) * g.its)), z);
              const float4 wr = *reinterpret_cast<const float4*>(sw + (kt * 9 + kh * 3 + kw) * PD + c0);
              v[0] = fmaf(z[0], wr.x, v[0]); v[1] = fmaf(z[1], wr.y, v[1]);
              v[2] = fmaf(z[2], wr.z, v[2]); v[3] = fmaf(z[3], wr.w, v[3]);
            }
          }
        }
      }
    }
    if (tok >= 1) {  // advance (to, ho, wo) to this warp's next token
      wo += step_w;
      if (wo >= g.Wo) { wo -= g.Wo; ++ho; }
      ho += step_h;
      if (ho >= g.Ho) { ho -= g.Ho; ++to; }
      to += step_t;
    }
    const float mean = warp_sum(v[0] + v[1] + v[2] + v[3]) * (1.f / PD);
    float xh[4], q = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      xh[j] = act ? v[j] - mean : 0.f;
      q += xh[j] * xh[j];
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / PD) + eps);
    const float gmv[4] = {gm.x, gm.y, gm.z, gm.w};
    float gy[4], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      xh[j] *= rstd;
      gy[j] = dy[j] * gmv[j];
      s1 += gy[j];
      s2 += gy[j] * xh[j];
      ag[j] += dy[j] * xh[j];
      ab[j] += dy[j];
    }
    s1 = warp_sum(s1) * (1.f / PD);
    s2 = warp_sum(s2) * (1.f / PD);
    if (act) {
      float dp[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) dp[j] = rstd * (gy[j] - s1 - xh[j] * s2);
      *reinterpret_cast<uint2*>(dpre + i * PD + c0) = pack4(dp);
    }
  }
  if (act) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      atomicAdd(&sred[c0 + j], ag[j]);
      atomicAdd(&sred[PD + c0 + j], ab[j]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < PD; c += blockDim.x) {
    atomicAdd(&dgamma[c], sred[c]);
    atomicAdd(&dbeta[c], sred[PD + c]);
  }
}

// LayerNorm backward on the SAVED pre-LayerNorm rows (svit_pool_ln_fwd_save): pure streaming, 16 lanes per output token
// (12 active, 8 channels each, 16-byte accesses), two tokens per warp instruction, reductions over 16 lanes.
__device__ __forceinline__ float group16_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, 16);
  return v;
}
__device__ __forceinline__ void unpack8p(const uint4 v, float f[8]) {
  f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
  f[4] = __uint_as_float(v.z << 16); f[5] = __uint_as_float(v.z & 0xffff0000u);
  f[6] = __uint_as_float(v.w << 16); f[7] = __uint_as_float(v.w & 0xffff0000u);
}

__global__ void __launch_bounds__(256) pool_bwd_pre_saved_kernel(const bf16* __restrict__ pre, const float* __restrict__ gamma,
                                                                 const bf16* __restrict__ dout, bf16* __restrict__ dpre,
                                                                 float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                 int64_t rows, float eps) {
  __shared__ float sred[16][2 * PD + 1];  // one row of partial (dgamma | dbeta) per 16-lane token group: no shared atomics
  const int sub = threadIdx.x & 15;
  const bool act = sub < PD / 8;
  const int c0 = act ? sub * 8 : 0;
  float gm[8];
  {
    const float4 ga = *reinterpret_cast<const float4*>(gamma + c0), gb = *reinterpret_cast<const float4*>(gamma + c0 + 4);
    gm[0] = ga.x; gm[1] = ga.y; gm[2] = ga.z; gm[3] = ga.w; gm[4] = gb.x; gm[5] = gb.y; gm[6] = gb.z; gm[7] = gb.w;
  }
  float ag[8], ab[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) ag[j] = ab[j] = 0.f;
  const int64_t gpb = blockDim.x >> 4;  // token groups per CTA
  const int64_t iters = (rows + gridDim.x * gpb - 1) / (gridDim.x * gpb);  // uniform trip count: full-mask shuffles
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t i = (it * gridDim.x + blockIdx.x) * gpb + (threadIdx.x >> 4);
    const bool ok = act && i < rows;
    float v[8], dy[8];
    unpack8p(ok ? __ldg(reinterpret_cast<const uint4*>(pre + i * PD + c0)) : make_uint4(0, 0, 0, 0), v);
    unpack8p(ok ? __ldg(reinterpret_cast<const uint4*>(dout + i * PD + c0)) : make_uint4(0, 0, 0, 0), dy);
    float sm = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) sm += v[j];
    const float mean = group16_sum(sm) * (1.f / PD);
    float xh[8], q = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      xh[j] = act ? v[j] - mean : 0.f;
      q = fmaf(xh[j], xh[j], q);
    }
    const float rstd = rsqrtf(group16_sum(q) * (1.f / PD) + eps);
    float gy[8], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      xh[j] *= rstd;
      gy[j] = dy[j] * gm[j];
      s1 += gy[j];
      s2 = fmaf(gy[j], xh[j], s2);
      ag[j] = fmaf(dy[j], xh[j], ag[j]);
      ab[j] += dy[j];
    }
    s1 = group16_sum(s1) * (1.f / PD);
    s2 = group16_sum(s2) * (1.f / PD);
    if (ok) {
      float dp[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) dp[j] = rstd * (gy[j] - s1 - xh[j] * s2);
      uint4 o;
      o.x = pack4(dp).x; o.y = pack4(dp).y; o.z = pack4(dp + 4).x; o.w = pack4(dp + 4).y;
      *reinterpret_cast<uint4*>(dpre + i * PD + c0) = o;
    }
  }
  if (act) {
    float* mine = sred[threadIdx.x >> 4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mine[c0 + j] = ag[j];
      mine[PD + c0 + j] = ab[j];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * PD; c += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int gidx = 0; gidx < 16; ++gidx) t += sred[gidx][c];
    atomicAdd(c < PD ? &dgamma[c] : &dbeta[c - PD], t);
  }
}

// block = 288 threads: quad = tid % 24 (channels 4 quad .. 4 quad + 3, 8-byte accesses), tg = (tid / 24) % 3 (temporal tap
// plane), part = tid / 72 (token index mod 4 inside the chunk).  grid = (chunks, B * h).
__global__ void __launch_bounds__(288, 3) pool_bwd_dw_kernel(const bf16* __restrict__ in, Geom g,
                                                          const float* __restrict__ frac, const bf16* __restrict__ dpre,
                                                          float* __restrict__ dw, int chunk) {
  __shared__ float sacc[3 * 72 * 36];  // partials of parts 1..3, later the combined [96][27] gradient
  __shared__ float swe[4 * 24 * 4];
  const int quad = threadIdx.x % 24, tg = (threadIdx.x / 24) % 3, part = threadIdx.x / 72;
  const int Lo = g.T * g.Ho * g.Wo, L = g.T * g.H * g.W;
  const int Nout = 1 + Lo + g.O;
  const int bh = blockIdx.y, head = bh % g.h, b = bh / g.h;
  const bf16* zin = in + b * g.in_bs + head * g.in_hs + 4 * quad;
  const bf16* dp_base = dpre + (int64_t)bh * Nout * PD + 4 * quad;
  const int t0 = blockIdx.x * chunk;
  const int t1 = t0 + chunk < Nout ? t0 + chunk : Nout;
  float acc[9][4], aweff[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f;
  // (to, ho, wo) of the thread's current patch token, advanced by 4 per iteration (no division in the loop)
  int wo, ho, to;
  {
    const int first = t0 + part < 1 ? t0 + part + 4 : t0 + part;  // first token this thread decodes as a patch token
    const unsigned p = first - 1;
    const unsigned pr = p / (unsigned)g.Wo;
    wo = (int)(p - pr * g.Wo); to = (int)(pr / (unsigned)g.Ho); ho = (int)(pr - to * g.Ho);
  }
#pragma unroll 2
  for (int tok = t0 + part; tok < t1; tok += 4) {
    if (tok == 0) continue;  // cls passes through: no weight gradient
    float dp[4];
    unpack4(*reinterpret_cast<const uint2*>(dp_base + tok * PD), dp);
    if (tok > Lo) {
      if (tg == 0) {
        float z[4];
        unpack4(*reinterpret_cast<const uint2*>(zin + (tok - Lo + L) * g.its), z);
#pragma unroll
        for (int e = 0; e < 4; ++e) aweff[e] = fmaf(dp[e], z[e], aweff[e]);
      }
      continue;
    }
    const int cwo = wo, cho = ho, cto = to;
    wo += 4;
    while (wo >= g.Wo) {
      wo -= g.Wo;
      if (++ho == g.Ho) { ho = 0; ++to; }
    }
    const int t = cto - 1 + tg;
    if (t < 0 || t >= g.T) continue;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int hh = cho * g.s - 1 + kh;
      if (hh < 0 || hh >= g.H) continue;
      const int rowbase = 1 + (t * g.H + hh) * g.W;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int ww = cwo * g.s - 1 + kw;
        if (ww < 0 || ww >= g.W) continue;
        float z[4];
        unpack4(*reinterpret_cast<const uint2*>(zin + (rowbase + ww) * g.its), z);
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[kh * 3 + kw][e] = fmaf(dp[e], z[e], acc[kh * 3 + kw][e]);
      }
    }
  }
  // object-token path: every tap of the plane gets tap_frac * d w_eff (tg 0 holds d w_eff; share it through smem)
  if (tg == 0) {
#pragma unroll
    for (int e = 0; e < 4; ++e) swe[(part * 24 + quad) * 4 + e] = aweff[e];
  }
  if (part > 0) {
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
      for (int e = 0; e < 4; ++e) sacc[((part - 1) * 72 + tg * 24 + quad) * 36 + k * 4 + e] = acc[k][e];
  }
  __syncthreads();
  if (part == 0) {
    float e4[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
      e4[e] = swe[quad * 4 + e] + swe[(24 + quad) * 4 + e] + swe[(48 + quad) * 4 + e] + swe[(72 + quad) * 4 + e];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float f = frac[tg * 9 + k];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v = fmaf(f, e4[e], acc[k][e]);
#pragma unroll
        for (int pp = 0; pp < 3; ++pp) v += sacc[(pp * 72 + tg * 24 + quad) * 36 + k * 4 + e];
        acc[k][e] = v;
      }
    }
  }
  __syncthreads();
  // laid out like dw ([channel][tap]) so that the global atomics of a warp fall on consecutive addresses (one L2
  // transaction per 128-byte line instead of one per element)
  float* sdw = sacc;  // [96 * 27]
  if (part == 0) {
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
      for (int e = 0; e < 4; ++e) sdw[(4 * quad + e) * TAPS + tg * 9 + k] = acc[k][e];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < PD * TAPS; i += blockDim.x) atomicAdd(&dw[i], sdw[i]);
}

// Strip form of the weight-gradient kernel (Wo % 7 == 0, stride 1 / 2 / 4): a thread (channel quad, temporal tap plane,
// part) takes 7 consecutive output tokens of one output row at a time.  Their 7 dpre quads stay in registers and every
// input token of the three rows under the strip is loaded ONCE and multiplied into all the outputs whose window covers
// it (9 / 15 / 21 loads per row for 84 multiply-adds, where the token form issues 21 loads), with every tap index
// resolved at compile time.  The loop of the token form was bound by its loads (10 per 36 FMAs).
template <int S>
__global__ void __launch_bounds__(288, 2) pool_bwd_dw_strip_kernel(const bf16* __restrict__ in, Geom g,
                                                                   const float* __restrict__ frac,
                                                                   const bf16* __restrict__ dpre, float* __restrict__ dw,
                                                                   int strips_per_cta) {
  constexpr int SW = 7, NI = (SW - 1) * S + 3;
  __shared__ float sacc[3 * 72 * 36];
  __shared__ float swe[4 * 24 * 4];
  const int quad = threadIdx.x % 24, tg = (threadIdx.x / 24) % 3, part = threadIdx.x / 72;
  const int Lo = g.T * g.Ho * g.Wo, L = g.T * g.H * g.W;
  const int Nout = 1 + Lo + g.O;
  const int bh = blockIdx.y, head = bh % g.h, b = bh / g.h;
  const bf16* zin = in + b * g.in_bs + head * g.in_hs + 4 * quad;
  const bf16* dp_base = dpre + (int64_t)bh * Nout * PD + 4 * quad;
  float acc[9][4], aweff[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f;
  const int spr = g.Wo / SW;  // strips per output row
  const int nstrips = g.T * g.Ho * spr;
  const int st0 = blockIdx.x * strips_per_cta;
  const int st1 = st0 + strips_per_cta < nstrips ? st0 + strips_per_cta : nstrips;
  for (int st = st0 + part; st < st1; st += 4) {
    const int sx = st % spr, r = st / spr;
    const int ho = r % g.Ho, to = r / g.Ho;
    const int t = to - 1 + tg;
    if (t < 0 || t >= g.T) continue;
    const int wo0 = sx * SW;
    float dp[SW][4];
    {
      const bf16* dps = dp_base + (1 + (to * g.Ho + ho) * g.Wo + wo0) * PD;
#pragma unroll
      for (int o = 0; o < SW; ++o) unpack4(__ldg(reinterpret_cast<const uint2*>(dps + o * PD)), dp[o]);
    }
    const int wbase = wo0 * S - 1;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int hh = ho * S - 1 + kh;
      if (hh < 0 || hh >= g.H) continue;
      const bf16* zrow = zin + (1 + (t * g.H + hh) * g.W + wbase) * g.its;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        if (i % S > 2 && S > 3) continue;  // stride 4: every fourth input column lies between the windows
        const int ww = wbase + i;
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        if (ww >= 0 && ww < g.W) unpack4(__ldg(reinterpret_cast<const uint2*>(zrow + i * g.its)), z);
#pragma unroll
        for (int o = 0; o < SW; ++o) {
          const int kw = i - S * o;
          if (kw >= 0 && kw <= 2) {
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[kh * 3 + kw][e] = fmaf(dp[o][e], z[e], acc[kh * 3 + kw][e]);
          }
        }
      }
    }
  }
  // object tokens (no window: z * w_eff): the first CTA of the slice takes them
  if (blockIdx.x == 0 && tg == 0) {
    for (int tok = Lo + 1 + part; tok < Nout; tok += 4) {
      float dpo[4], z[4];
      unpack4(__ldg(reinterpret_cast<const uint2*>(dp_base + tok * PD)), dpo);
      unpack4(__ldg(reinterpret_cast<const uint2*>(zin + (tok - Lo + L) * g.its)), z);
#pragma unroll
      for (int e = 0; e < 4; ++e) aweff[e] = fmaf(dpo[e], z[e], aweff[e]);
    }
  }
  if (tg == 0) {
#pragma unroll
    for (int e = 0; e < 4; ++e) swe[(part * 24 + quad) * 4 + e] = aweff[e];
  }
  if (part > 0) {
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
      for (int e = 0; e < 4; ++e) sacc[((part - 1) * 72 + tg * 24 + quad) * 36 + k * 4 + e] = acc[k][e];
  }
  __syncthreads();
  if (part == 0) {
    float e4[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
      e4[e] = swe[quad * 4 + e] + swe[(24 + quad) * 4 + e] + swe[(48 + quad) * 4 + e] + swe[(72 + quad) * 4 + e];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float f = frac[tg * 9 + k];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v = fmaf(f, e4[e], acc[k][e]);
#pragma unroll
        for (int pp = 0; pp < 3; ++pp) v += sacc[(pp * 72 + tg * 24 + quad) * 36 + k * 4 + e];
        acc[k][e] = v;
      }
    }
  }
  __syncthreads();
  float* sdw = sacc;  // [96 * 27], laid out like dw: coalesced global atomics
  if (part == 0) {
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
      for (int e = 0; e < 4; ++e) sdw[(4 * quad + e) * TAPS + tg * 9 + k] = acc[k][e];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < PD * TAPS; i += blockDim.x) atomicAdd(&dw[i], sdw[i]);
}

__device__ __forceinline__ void unpack8w(const uint4 v, float f[8]) {
  f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
  f[4] = __uint_as_float(v.z << 16); f[5] = __uint_as_float(v.z & 0xffff0000u);
  f[6] = __uint_as_float(v.w << 16); f[7] = __uint_as_float(v.w & 0xffff0000u);
}

// 16 lanes per INPUT token, 12 of them active with 8 channels each (16-byte accesses; two tokens per warp instruction).
// V8 = false keeps the 8-byte form (24 lanes x 4 channels) for slices that are only 8-byte aligned.
template <int LS, bool V8>  // LS = log2(stride) for power-of-two strides, -1 = generic
__global__ void __launch_bounds__(256) pool_bwd_in_kernel(const bf16* __restrict__ dpre, Geom g,
                                                          const float* __restrict__ w, const float* __restrict__ frac,
                                                          bf16* __restrict__ dz) {
  __shared__ __align__(16) float sw[TAPS * PD];
  __shared__ __align__(16) float sweff[PD];
  load_weights(w, frac, sw, sweff);
  constexpr int LPT = V8 ? 16 : 32;       // lanes per token
  constexpr int NCH = V8 ? 8 : 4;         // channels per lane
  const int sub = threadIdx.x & (LPT - 1);
  if (sub >= PD / NCH) return;
  const int c0 = sub * NCH;
  const int Lo = g.T * g.Ho * g.Wo, L = g.T * g.H * g.W;
  const int Nout = 1 + Lo + g.O, Nin = 1 + L + g.O;
  const int tpb = blockDim.x / LPT;       // tokens per CTA pass
  const int bh = blockIdx.y, head = bh % g.h, b = bh / g.h;
  const bf16* dp = dpre + (int64_t)bh * Nout * PD + c0;
  bf16* dzb = dz + b * g.in_bs + head * g.in_hs + c0;
  const int smask = (1 << (LS >= 0 ? LS : 0)) - 1;  // unused by the generic (LS < 0) instantiation
  const int step = gridDim.x * tpb;
  const int step_w = step % g.W, step_h = (step / g.W) % g.H, step_t = step / (g.W * g.H);
  int ww, hh, t;
  {
    const int tok0 = blockIdx.x * tpb + threadIdx.x / LPT;
    const unsigned p = (tok0 < 1 ? tok0 + step : tok0) - 1;
    const unsigned pr = p / (unsigned)g.W;
    ww = (int)(p - pr * g.W); t = (int)(pr / (unsigned)g.H); hh = (int)(pr - t * g.H);
  }
  auto load = [&](const bf16* src, float f[NCH]) {
    if (V8) {
      float t8[8];
      unpack8w(__ldg(reinterpret_cast<const uint4*>(src)), t8);
#pragma unroll
      for (int e = 0; e < NCH; ++e) f[e] = t8[e];
    } else {
      float t4[4];
      unpack4(__ldg(reinterpret_cast<const uint2*>(src)), t4);
#pragma unroll
      for (int e = 0; e < NCH; ++e) f[e] = t4[e];
    }
  };
  for (int tok = blockIdx.x * tpb + threadIdx.x / LPT; tok < Nin; tok += step) {
    float v[NCH];
#pragma unroll
    for (int e = 0; e < NCH; ++e) v[e] = 0.f;
    if (tok == 0) {
      load(dp, v);
    } else if (tok > L) {
      float z[NCH];
      load(dp + (tok - L + Lo) * PD, z);
#pragma unroll
      for (int e = 0; e < NCH; ++e) v[e] = z[e] * sweff[c0 + e];
    } else {
#pragma unroll
      for (int kt = 0; kt < 3; ++kt) {
        const int to = t + 1 - kt;
        if (to < 0 || to >= g.T) continue;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const int num = hh + 1 - kh;
          if (num < 0 || (LS >= 0 ? (num & smask) != 0 : num % g.s != 0)) continue;
          const int ho = LS >= 0 ? num >> LS : num / g.s;
          if (ho >= g.Ho) continue;
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const int numw = ww + 1 - kw;
            if (numw < 0 || (LS >= 0 ? (numw & smask) != 0 : numw % g.s != 0)) continue;
            const int wo = LS >= 0 ? numw >> LS : numw / g.s;
            if (wo >= g.Wo) continue;
            float z[NCH];
            load(dp + (1 + (to * g.Ho + ho) * g.Wo + wo) * PD, z);
            const float* wr = sw + (kt * 9 + kh * 3 + kw) * PD + c0;
#pragma unroll
            for (int e = 0; e < NCH; e += 4) {
              const float4 w4 = *reinterpret_cast<const float4*>(wr + e);
              v[e] = fmaf(z[e], w4.x, v[e]); v[e + 1] = fmaf(z[e + 1], w4.y, v[e + 1]);
              v[e + 2] = fmaf(z[e + 2], w4.z, v[e + 2]); v[e + 3] = fmaf(z[e + 3], w4.w, v[e + 3]);
            }
          }
        }
      }
    }
    if (V8) {
      uint4 o;
      o.x = pack4(v).x; o.y = pack4(v).y; o.z = pack4(v + 4).x; o.w = pack4(v + 4).y;
      *reinterpret_cast<uint4*>(dzb + tok * g.its) = o;
    } else {
      *reinterpret_cast<uint2*>(dzb + tok * g.its) = pack4(v);
    }
    if (tok >= 1) {
      ww += step_w;
      if (ww >= g.W) { ww -= g.W; ++hh; }
      hh += step_h;
      if (hh >= g.H) { hh -= g.H; ++t; }
      t += step_t;
    }
  }
}

// Strip form of the input-gradient kernel (transposed depthwise conv): a lane group takes 7 (stride 1, 8 channels per
// lane) or 14 (stride 2, 4 channels per lane) consecutive input tokens of one row.  Per (temporal, vertical) tap pair it
// loads the 9 / 8 dpre tokens under the strip once and the 3 x NCH weights once, and every (input, tap) product is
// resolved at compile time -- the token form issues one 16-byte load and two shared-memory weight loads per 8 FMAs.
template <int S>
__global__ void __launch_bounds__(256, 2) pool_bwd_in_strip_kernel(const bf16* __restrict__ dpre, Geom g,
                                                                   const float* __restrict__ w,
                                                                   const float* __restrict__ frac, bf16* __restrict__ dz) {
  __shared__ __align__(16) float sw[TAPS * PD];
  __shared__ __align__(16) float sweff[PD];
  load_weights(w, frac, sw, sweff);
  constexpr int NCH = S == 1 ? 8 : 4, SWI = S == 1 ? 7 : 14, LPT = S == 1 ? 16 : 32, NO = S == 1 ? 9 : 8;
  constexpr int OFF = S == 1 ? 1 : 0;
  const int sub = threadIdx.x & (LPT - 1);
  if (sub >= PD / NCH) return;
  const int c0 = sub * NCH;
  const int Lo = g.T * g.Ho * g.Wo, L = g.T * g.H * g.W;
  const int Nout = 1 + Lo + g.O;
  const int bh = blockIdx.y, head = bh % g.h, b = bh / g.h;
  const bf16* dp = dpre + (int64_t)bh * Nout * PD + c0;
  bf16* dzb = dz + b * g.in_bs + head * g.in_hs + c0;
  auto load = [&](const bf16* src, float f[NCH]) {
    if (NCH == 8) {
      float t8[8];
      unpack8w(__ldg(reinterpret_cast<const uint4*>(src)), t8);
#pragma unroll
      for (int e = 0; e < NCH; ++e) f[e] = t8[e];
    } else {
      float t4[4];
      unpack4(__ldg(reinterpret_cast<const uint2*>(src)), t4);
#pragma unroll
      for (int e = 0; e < NCH; ++e) f[e] = t4[e];
    }
  };
  auto store = [&](bf16* dst, const float f[NCH]) {
    if (NCH == 8) {
      uint4 o;
      o.x = pack4(f).x; o.y = pack4(f).y; o.z = pack4(f + 4).x; o.w = pack4(f + 4).y;
      *reinterpret_cast<uint4*>(dst) = o;
    } else {
      *reinterpret_cast<uint2*>(dst) = pack4(f);
    }
  };
  const int spr = g.W / SWI;
  const int nstrips = g.T * g.H * spr;
  const int spb = blockDim.x / LPT;
  for (int st = blockIdx.x * spb + threadIdx.x / LPT; st < nstrips; st += gridDim.x * spb) {
    const int sx = st % spr, r = st / spr;
    const int hh = r % g.H, t = r / g.H;
    const int ww0 = sx * SWI;
    const int wo_base = S == 1 ? ww0 - 1 : ww0 >> 1;
    float v[SWI][NCH];
#pragma unroll
    for (int j = 0; j < SWI; ++j)
#pragma unroll
      for (int e = 0; e < NCH; ++e) v[j][e] = 0.f;
#pragma unroll
    for (int kt = 0; kt < 3; ++kt) {
      const int to = t + 1 - kt;
      if (to < 0 || to >= g.T) continue;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int num = hh + 1 - kh;
        if (num < 0 || (S == 2 && (num & 1))) continue;
        const int ho = S == 2 ? num >> 1 : num;
        if (ho >= g.Ho) continue;
        const bf16* zr = dp + (1 + (to * g.Ho + ho) * g.Wo + wo_base) * PD;
        float wk[3][NCH];
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
#pragma unroll
          for (int e = 0; e < NCH; e += 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(sw + (kt * 9 + kh * 3 + kw) * PD + c0 + e);
            wk[kw][e] = w4.x; wk[kw][e + 1] = w4.y; wk[kw][e + 2] = w4.z; wk[kw][e + 3] = w4.w;
          }
#pragma unroll
        for (int i = 0; i < NO; ++i) {
          const int wo = wo_base + i;
          float z[NCH];
#pragma unroll
          for (int e = 0; e < NCH; ++e) z[e] = 0.f;
          if (wo >= 0 && wo < g.Wo) load(zr + i * PD, z);
#pragma unroll
          for (int j = 0; j < SWI; ++j) {
            const int kw = j + 1 + OFF - S * i;  // input ww0 + j reads output wo through tap kw = ww + 1 - S wo
            if (kw >= 0 && kw <= 2) {
#pragma unroll
              for (int e = 0; e < NCH; ++e) v[j][e] = fmaf(z[e], wk[kw][e], v[j][e]);
            }
          }
        }
      }
    }
    bf16* dst = dzb + (1 + (t * g.H + hh) * g.W + ww0) * g.its;
#pragma unroll
    for (int j = 0; j < SWI; ++j) store(dst + j * g.its, v[j]);
  }
  if (blockIdx.x == 0) {  // cls row (pass-through) and the object rows (z * w_eff)
    for (int n = threadIdx.x / LPT; n <= g.O; n += spb) {
      float z[NCH];
      if (n == 0) {
        load(dp, z);
        store(dzb, z);
      } else {
        load(dp + (Lo + n) * PD, z);
#pragma unroll
        for (int e = 0; e < NCH; ++e) z[e] *= sweff[c0 + e];
        store(dzb + (L + n) * g.its, z);
      }
    }
  }
}

}  // namespace

int svit_pool_ln_bwd_bf16_supported(const void* in, int64_t in_bs, int64_t in_ts, int64_t in_hs, const void* dout,
                                    const void* dpre, const void* dz) {
  if (in_bs % 4 || in_ts % 4 || in_hs % 4) return 0;
  const uintptr_t m = reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(dout) |
                      reinterpret_cast<uintptr_t>(dpre) | reinterpret_cast<uintptr_t>(dz);
  return (m & 7) == 0;
}

int svit_pool_ln_bwd_bf16(const void* in, int64_t in_bs, int64_t in_ts, int64_t in_hs, const float* conv_w,
                          const float* tap_frac, const float* gamma, const void* dout, void* dpre, void* dz, float* dw,
                          float* dgamma, float* dbeta, int B, int h, int T, int H, int W, int O, int s, float eps,
                          cudaStream_t st, const void* pre) {
  Geom g;
  g.B = B; g.h = h; g.T = T; g.H = H; g.W = W; g.O = O; g.s = s;
  g.Ho = (H - 1) / s + 1;
  g.Wo = (W - 1) / s + 1;
  g.in_bs = in_bs; g.in_ts = in_ts; g.in_hs = in_hs; g.its = (int)in_ts;
  const int64_t Nout = 1 + (int64_t)T * g.Ho * g.Wo + O, Nin = 1 + (int64_t)T * H * W + O;
  const int64_t tok_out = (int64_t)B * h * Nout, tok_in = (int64_t)B * h * Nin;
  if (B * h > 65535 || Nin * in_ts >= (1ll << 31) || Nin * PD >= (1ll << 31)) return SVIT_ENOTSUP;  // 32-bit offsets
  (void)tok_in;
  const unsigned BH = (unsigned)(B * h);
  auto gx = [&](int64_t n) {  // CTAs along x so that x * BH is about 8 CTAs per SM, 8 tokens per CTA pass
    int64_t want = ceil_div64((int64_t)svit_num_sms() * 8, BH), need = ceil_div64(n, 8);
    return (unsigned)(need < want ? need : want);
  };
  if (pre && ((reinterpret_cast<uintptr_t>(pre) | reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(dpre)) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(gamma) & 15) == 0) {
    // every CTA ends with 192 global atomics on the same two cache lines (~10 ns each, serialised in L2): at least four
    // 16-token passes per CTA and at most 4 CTAs per SM keep that tail at a few microseconds
    int64_t ctas = ceil_div64(tok_out, 64);
    const int64_t cap = (int64_t)svit_num_sms() * 4;
    if (ctas > cap) ctas = cap;
    pool_bwd_pre_saved_kernel<<<(unsigned)ctas, 256, 0, st>>>((const bf16*)pre, gamma, (const bf16*)dout, (bf16*)dpre, dgamma,
                                                             dbeta, tok_out, eps);
  } else {
    pool_bwd_pre_kernel<<<dim3(gx(Nout), BH), 256, 0, st>>>((const bf16*)in, g, conv_w, tap_frac, gamma,
                                                                  (const bf16*)dout, (bf16*)dpre, dgamma, dbeta, eps,
                                                                  (const bf16*)pre);
  }
  SVIT_CHECK_LAUNCH();
  int64_t chunk = ceil_div64(tok_out, (int64_t)svit_num_sms() * 8);  // ~8 CTAs (72 warps) per SM: latency-bound loop
  if (chunk < 32) chunk = 32;
  if (chunk > 2048) chunk = 2048;
  if (g.Wo % 7 == 0 && (s == 1 || s == 2 || s == 4)) {
    const int64_t nstrips = (int64_t)T * g.Ho * (g.Wo / 7);
    int64_t spc = ceil_div64(nstrips * BH, (int64_t)svit_num_sms() * 4);  // ~4 CTAs per SM over the whole launch
    if (spc < 8) spc = 8;
    if (spc > nstrips) spc = nstrips;
    const dim3 gs((unsigned)ceil_div64(nstrips, spc), BH);
    if (s == 1) pool_bwd_dw_strip_kernel<1><<<gs, 288, 0, st>>>((const bf16*)in, g, tap_frac, (const bf16*)dpre, dw, (int)spc);
    else if (s == 2) pool_bwd_dw_strip_kernel<2><<<gs, 288, 0, st>>>((const bf16*)in, g, tap_frac, (const bf16*)dpre, dw, (int)spc);
    else pool_bwd_dw_strip_kernel<4><<<gs, 288, 0, st>>>((const bf16*)in, g, tap_frac, (const bf16*)dpre, dw, (int)spc);
  } else {
    pool_bwd_dw_kernel<<<dim3((unsigned)ceil_div64(Nout, chunk), (unsigned)(B * h)), 288, 0, st>>>(
        (const bf16*)in, g, tap_frac, (const bf16*)dpre, dw, (int)chunk);
  }
  SVIT_CHECK_LAUNCH();
  // 16-byte form when the packed qkv slice and the scratch rows are 16-byte aligned (always for the model's layouts)
  const bool v8 = in_bs % 8 == 0 && in_ts % 8 == 0 && in_hs % 8 == 0 &&
                  ((reinterpret_cast<uintptr_t>(dpre) | reinterpret_cast<uintptr_t>(dz)) & 15) == 0;
  auto gin_x = [&](int64_t n, int tpb) {
    int64_t want = ceil_div64((int64_t)svit_num_sms() * 8, BH), need = ceil_div64(n, tpb);
    return (unsigned)(need < want ? need : want);
  };
#define IN_LAUNCH(LS)                                                                                                         \
  if (v8) pool_bwd_in_kernel<LS, true><<<dim3(gin_x(Nin, 16), BH), 256, 0, st>>>((const bf16*)dpre, g, conv_w, tap_frac, (bf16*)dz); \
  else pool_bwd_in_kernel<LS, false><<<dim3(gin_x(Nin, 8), BH), 256, 0, st>>>((const bf16*)dpre, g, conv_w, tap_frac, (bf16*)dz);
  auto strip_grid = [&](int64_t nstrips, int per_pass) {
    int64_t want = ceil_div64((int64_t)svit_num_sms() * 2, BH), need = ceil_div64(nstrips, per_pass);
    return dim3((unsigned)(need < want ? need : want), BH);
  };
  if (s == 1 && v8 && W % 7 == 0) {
    pool_bwd_in_strip_kernel<1><<<strip_grid((int64_t)T * H * (W / 7), 16), 256, 0, st>>>((const bf16*)dpre, g, conv_w, tap_frac,
                                                                                      (bf16*)dz);
    SVIT_CHECK_LAUNCH();
    return 0;
  }
  if (s == 2 && W % 14 == 0) {
    pool_bwd_in_strip_kernel<2><<<strip_grid((int64_t)T * H * (W / 14), 8), 256, 0, st>>>((const bf16*)dpre, g, conv_w, tap_frac,
                                                                                      (bf16*)dz);
    SVIT_CHECK_LAUNCH();
    return 0;
  }
  switch (s) {
    case 1: IN_LAUNCH(0) break;
    case 2: IN_LAUNCH(1) break;
    case 4: IN_LAUNCH(2) break;
    case 8: IN_LAUNCH(3) break;
    default: IN_LAUNCH(-1) break;
  }
#undef IN_LAUNCH
  SVIT_CHECK_LAUNCH();
  return 0;
}
