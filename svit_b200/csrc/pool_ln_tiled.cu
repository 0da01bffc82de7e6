// attention_pool forward (conv + object-token scale + LayerNorm), bf16 production kernels.
// Reference: slowfast/models/attention.py:13-65.  HBM-bound: every input token slice is read once from
// DRAM (halo re-reads are L2 hits) and every output token written once.
//
// Mapping shared by both kernels: a HALF-WARP owns one output token; lane l16 owns the six channels
// {2w, 2w+1 : w = l16, l16+16, l16+32} as three bf16x2 words, so a token slice (192 B) is three conflict-free
// 64-byte half-warp accesses and LayerNorm(96) is a 4-step xor-shuffle inside the half-warp.
//
//   pool_ln_s1_tiled_kernel   stride (1,1,1): CTA = 8 x 14 output tile marched over T with a 3-plane rolling
//                             window of (8+2) x (14+2) token slices in shared memory (cp.async, zero fill);
//                             each half-warp produces a strip of 7 outputs with a sliding 3x9 register window,
//                             so one shared-memory word feeds up to 3 x 6 FMAs.
//   pool_ln_direct_kernel     stride (1,s,s), s >= 2 (windows barely overlap): taps straight from global / L2;
//                             also emits the cls and object-token rows for both kernels.
#include "common.cuh"

namespace {

constexpr int PD = 96;
constexpr int TAPS = 27;
constexpr int TW = 14, TH = 8, STRIP = 7;
constexpr int PW = TW + 2, PH = TH + 2;            // plane with halo
constexpr int PLANE_WORDS = PH * PW * (PD / 2);    // bf16x2 words
constexpr int SMEM_TILED = 3 * PLANE_WORDS * 4 + TAPS * PD * 4;

struct Geom {
  int B, h, T, H, W, Ho, Wo, O, s;
  int64_t in_bs, in_ts, in_hs;
};

__device__ __forceinline__ float lo_f(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float hi_f(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// reduction inside one half-warp; the mask names only this half so the two halves of a warp may diverge
__device__ __forceinline__ float half_sum(float v) {
  const unsigned mask = 0xffffu << (threadIdx.x & 16);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}

// LayerNorm of one token held as 6 floats per lane across a half-warp, then store as 3 bf16x2 words.
__device__ __forceinline__ void ln_store(const float v[6], const float g[6], const float b[6], float eps,
                                         uint32_t* __restrict__ dst, int l16) {
  const float mean = half_sum(v[0] + v[1] + v[2] + v[3] + v[4] + v[5]) * (1.f / PD);
  float d[6], q = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    d[i] = v[i] - mean;
    q += d[i] * d[i];
  }
  const float rstd = rsqrtf(half_sum(q) * (1.f / PD) + eps);
#pragma unroll
  for (int j = 0; j < 3; ++j)
    dst[l16 + 16 * j] = pack2(d[2 * j] * rstd * g[2 * j] + b[2 * j], d[2 * j + 1] * rstd * g[2 * j + 1] + b[2 * j + 1]);
}

__device__ __forceinline__ void load_affine(const float* gamma, const float* beta, int l16, float g[6], float b[6]) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int c = 2 * (l16 + 16 * j);
    g[2 * j] = gamma[c]; g[2 * j + 1] = gamma[c + 1];
    b[2 * j] = beta[c]; b[2 * j + 1] = beta[c + 1];
  }
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 16 : 0;  // src-size 0 -> 16 bytes of zeros (conv padding)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------ stride 1
__global__ void __launch_bounds__(256, 2)
pool_ln_s1_tiled_kernel(const bf16* __restrict__ in, Geom g, const float* __restrict__ w, const float* __restrict__ gamma,
                        const float* __restrict__ beta, bf16* __restrict__ out, float eps) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint32_t* planes = reinterpret_cast<uint32_t*>(smem);                       // [3][PH][PW][48]
  float* sw = reinterpret_cast<float*>(smem + 3 * PLANE_WORDS * 4);           // [27][96]
  const int tiles_w = (g.W + TW - 1) / TW;
  const int w0 = (blockIdx.x % tiles_w) * TW, h0 = (blockIdx.x / tiles_w) * TH;
  const int head = blockIdx.y, b = blockIdx.z;
  const bf16* zin = in + (int64_t)b * g.in_bs + (int64_t)head * g.in_hs;
  const int64_t Nout = 1 + (int64_t)g.T * g.H * g.W + g.O;
  bf16* obase = out + ((int64_t)b * g.h + head) * Nout * PD;

  for (int i = threadIdx.x; i < PD * TAPS; i += blockDim.x) sw[(i % TAPS) * PD + i / TAPS] = w[i];

  auto load_plane = [&](int t) {  // plane t -> slot (t + 3) % 3; planes -1 and T are the conv's zero padding
    uint32_t* dst = planes + ((t + 3) % 3) * PLANE_WORDS;
    const bool tv = t >= 0 && t < g.T;
    for (int i = threadIdx.x; i < PH * PW * 12; i += blockDim.x) {
      const int chunk = i % 12, pos = i / 12;
      const int pw = pos % PW, ph = pos / PW;
      const int hh = h0 - 1 + ph, ww = w0 - 1 + pw;
      const bool ok = tv && hh >= 0 && hh < g.H && ww >= 0 && ww < g.W;
      const bf16* src = ok ? zin + (1 + ((int64_t)t * g.H + hh) * g.W + ww) * g.in_ts + chunk * 8 : zin;
      cp_async16(dst + pos * 48 + chunk * 4, src, ok);
    }
  };
  load_plane(-1);
  load_plane(0);
  cp_async_commit();

  const int lane = threadIdx.x & 31, l16 = lane & 15;
  const int strip = (threadIdx.x >> 4);                 // 0..15 half-warps
  const int srow = strip >> 1, scol = (strip & 1) * STRIP;  // output row in tile, first output col in tile
  float gm[6], bt[6];
  load_affine(gamma, beta, l16, gm, bt);
  const int ho = h0 + srow;
  const bool row_ok = ho < g.H;

  for (int t = 0; t < g.T; ++t) {
    // plane t+1 -> slot (t+1)%3 (holds plane t-2, no longer needed)
    __syncthreads();  // everyone finished computing step t-1 (which read slot (t-2)%3 == (t+1)%3)
    load_plane(t + 1);
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();
    if (row_ok) {
      float acc[STRIP][6];
#pragma unroll
      for (int o = 0; o < STRIP; ++o)
#pragma unroll
        for (int c = 0; c < 6; ++c) acc[o][c] = 0.f;
#pragma unroll
      for (int kt = 0; kt < 3; ++kt) {
        const int tp = t - 1 + kt;
        const uint32_t* pl = planes + ((tp + 3) % 3) * PLANE_WORDS;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const uint32_t* rowp = pl + ((srow + kh) * PW + scol) * 48;
          float x[STRIP + 2][6];
#pragma unroll
          for (int p = 0; p < STRIP + 2; ++p)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const uint32_t wd = rowp[p * 48 + l16 + 16 * j];
              x[p][2 * j] = lo_f(wd);
              x[p][2 * j + 1] = hi_f(wd);
            }
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const float* wr = sw + ((kt * 3 + kh) * 3 + kw) * PD;
            float wt[6];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const float2 f = *reinterpret_cast<const float2*>(wr + 2 * (l16 + 16 * j));
              wt[2 * j] = f.x;
              wt[2 * j + 1] = f.y;
            }
#pragma unroll
            for (int o = 0; o < STRIP; ++o)
#pragma unroll
              for (int c = 0; c < 6; ++c) acc[o][c] = fmaf(x[o + kw][c], wt[c], acc[o][c]);
          }
        }
      }
#pragma unroll
      for (int o = 0; o < STRIP; ++o) {
        const int wo = w0 + scol + o;
        if (wo < g.W) {  // uniform across the half-warp
          uint32_t* dst = reinterpret_cast<uint32_t*>(obase + (1 + ((int64_t)t * g.H + ho) * g.W + wo) * PD);
          ln_store(acc[o], gm, bt, eps, dst, l16);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ direct
// One half-warp per output token.  mode 0: all tokens; mode 1: only cls + object tokens (companion of the tiled
// kernel, which writes the patch tokens).
__global__ void __launch_bounds__(256)
pool_ln_direct_kernel(const bf16* __restrict__ in, Geom g, const float* __restrict__ w, const float* __restrict__ frac,
                      const float* __restrict__ gamma, const float* __restrict__ beta, bf16* __restrict__ out, float eps,
                      int mode) {
  __shared__ float sw[TAPS * PD];
  __shared__ float sweff[PD];
  for (int i = threadIdx.x; i < PD * TAPS; i += blockDim.x) sw[(i % TAPS) * PD + i / TAPS] = w[i];
  __syncthreads();
  for (int c = threadIdx.x; c < PD; c += blockDim.x) {
    float a = 0.f;
    for (int t = 0; t < TAPS; ++t) a += sw[t * PD + c] * frac[t];
    sweff[c] = a;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, l16 = lane & 15;
  float gm[6], bt[6];
  load_affine(gamma, beta, l16, gm, bt);
  const int64_t Lo = (int64_t)g.T * g.Ho * g.Wo, L = (int64_t)g.T * g.H * g.W;
  const int64_t Nout = 1 + Lo + g.O;
  const int64_t per_bh = mode == 0 ? Nout : (int64_t)1 + g.O;
  const int64_t total = (int64_t)g.B * g.h * per_bh;
  const int64_t hw_per_grid = (int64_t)gridDim.x * (blockDim.x >> 4);
  for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 4) + (threadIdx.x >> 4); i < total; i += hw_per_grid) {
    const int64_t r = i % per_bh;
    const int64_t tok = mode == 0 ? r : (r == 0 ? 0 : Lo + r);
    const int head = (int)((i / per_bh) % g.h);
    const int b = (int)(i / (per_bh * g.h));
    const bf16* zin = in + (int64_t)b * g.in_bs + (int64_t)head * g.in_hs;
    float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (tok == 0 || tok > Lo) {
      const uint32_t* p = reinterpret_cast<const uint32_t*>(zin + (tok == 0 ? 0 : (tok - Lo + L)) * g.in_ts);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const uint32_t wd = p[l16 + 16 * j];
        const int c = 2 * (l16 + 16 * j);
        v[2 * j] = lo_f(wd) * (tok == 0 ? 1.f : sweff[c]);
        v[2 * j + 1] = hi_f(wd) * (tok == 0 ? 1.f : sweff[c + 1]);
      }
    } else {
      const int64_t pp = tok - 1;
      const int wo = (int)(pp % g.Wo), ho = (int)((pp / g.Wo) % g.Ho), to = (int)(pp / ((int64_t)g.Wo * g.Ho));
      // all 27 x 3 words of the window are requested before the first FMA (zero for padding taps), so the
      // DRAM / L2 latency is paid once per token, not once per tap
      uint32_t xw[TAPS][3];
#pragma unroll
      for (int kt = 0; kt < 3; ++kt) {
        const int t = to - 1 + kt;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const int hh = ho * g.s - 1 + kh;
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const int ww = wo * g.s - 1 + kw;
            const bool ok = t >= 0 && t < g.T && hh >= 0 && hh < g.H && ww >= 0 && ww < g.W;
            const uint32_t* q = reinterpret_cast<const uint32_t*>(
                zin + (ok ? (1 + ((int64_t)t * g.H + hh) * g.W + ww) * g.in_ts : 0));
#pragma unroll
            for (int j = 0; j < 3; ++j) xw[(kt * 3 + kh) * 3 + kw][j] = ok ? __ldg(q + l16 + 16 * j) : 0u;
          }
        }
      }
#pragma unroll
      for (int tap = 0; tap < TAPS; ++tap) {
        const float* wr = sw + tap * PD;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float2 f = *reinterpret_cast<const float2*>(wr + 2 * (l16 + 16 * j));
          v[2 * j] = fmaf(lo_f(xw[tap][j]), f.x, v[2 * j]);
          v[2 * j + 1] = fmaf(hi_f(xw[tap][j]), f.y, v[2 * j + 1]);
        }
      }
    }
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + (((int64_t)b * g.h + head) * Nout + tok) * PD);
    ln_store(v, gm, bt, eps, dst, l16);
  }
}

}  // namespace

// bf16 fast path of svit_pool_ln_fwd (pool_ln.cu dispatches here).  Requires 4-byte aligned token slices.
int svit_pool_ln_fwd_bf16(const void* in, int64_t in_bs, int64_t in_ts, int64_t in_hs, const float* conv_w,
                          const float* tap_frac, const float* gamma, const float* beta, void* out, int B, int h, int T,
                          int H, int W, int O, int s, float eps, cudaStream_t st) {
  Geom g;
  g.B = B; g.h = h; g.T = T; g.H = H; g.W = W; g.O = O; g.s = s;
  g.Ho = (H - 1) / s + 1; g.Wo = (W - 1) / s + 1;
  g.in_bs = in_bs; g.in_ts = in_ts; g.in_hs = in_hs;
  const int sms = svit_num_sms();
  const bool tiled = (s == 1) && (in_ts % 8 == 0) && (in_hs % 8 == 0) && (in_bs % 8 == 0) &&
                     ((reinterpret_cast<uintptr_t>(in) & 15) == 0) && H * W >= 49;
  if (tiled) {
    static bool configured = false;
    if (!configured) {
      SVIT_CUDA(cudaFuncSetAttribute(pool_ln_s1_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TILED));
      configured = true;
    }
    dim3 grid((unsigned)(((W + TW - 1) / TW) * ((H + TH - 1) / TH)), (unsigned)h, (unsigned)B);
    pool_ln_s1_tiled_kernel<<<grid, 256, SMEM_TILED, st>>>((const bf16*)in, g, conv_w, gamma, beta, (bf16*)out, eps);
    SVIT_CHECK_LAUNCH();
    const int64_t special = (int64_t)B * h * (1 + O);
    int blocks = (int)((special + 15) / 16);
    if (blocks > sms * 8) blocks = sms * 8;
    pool_ln_direct_kernel<<<blocks, 256, 0, st>>>((const bf16*)in, g, conv_w, tap_frac, gamma, beta, (bf16*)out, eps, 1);
    SVIT_CHECK_LAUNCH();
    return 0;
  }
  const int64_t total = (int64_t)B * h * (1 + (int64_t)T * g.Ho * g.Wo + O);
  int64_t blocks = (total + 15) / 16;
  if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
  if (blocks < 1) blocks = 1;
  pool_ln_direct_kernel<<<(unsigned)blocks, 256, 0, st>>>((const bf16*)in, g, conv_w, tap_frac, gamma, beta, (bf16*)out, eps, 0);
  SVIT_CHECK_LAUNCH();
  return 0;
}
