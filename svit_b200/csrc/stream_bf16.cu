// bf16 production variants of the HBM-bound streaming kernels (generic dtype versions live in elementwise.cu):
//   layernorm_bf16_kernel   8-byte (bf16x4) accesses; a group of C/12 lanes owns one token row
//   im2col_rows_kernel      PatchEmbed lowering (stem_helper.py:309-320): one CTA per (b, t', h') output row stages
//                           the Cin*kt*kh input rows it needs in shared memory once, then emits 16-byte column
//                           chunks of the [tokens, Kpad] matrix fully coalesced
#include "common.cuh"

namespace {

__device__ __forceinline__ void unpack4(uint2 v, float f[4]) {
  f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// G lanes per row, NQ bf16x4 quads per lane: C = 4 * G * NQ  (C = 96: G 8; 192: 16; 384: 32; 768: 32 with NQ 6)
template <int G, int NQ>
__global__ void __launch_bounds__(256) layernorm_bf16_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, bf16* __restrict__ y,
                                                             float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                             int64_t rows, float eps) {
  constexpr int C = 4 * G * NQ;
  const int lg = threadIdx.x % G;
  const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  const int64_t ngroups = (int64_t)gridDim.x * blockDim.x / G;
  float gm[NQ][4], bt[NQ][4];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const float4 g4 = *reinterpret_cast<const float4*>(gamma + 4 * (lg + G * q));
    const float4 b4 = *reinterpret_cast<const float4*>(beta + 4 * (lg + G * q));
    gm[q][0] = g4.x; gm[q][1] = g4.y; gm[q][2] = g4.z; gm[q][3] = g4.w;
    bt[q][0] = b4.x; bt[q][1] = b4.y; bt[q][2] = b4.z; bt[q][3] = b4.w;
  }
  // all lanes of a warp iterate the same number of times (shuffles need the full mask)
  const int64_t iters = (rows + ngroups - 1) / ngroups;
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t row = group + it * ngroups;
    const bool ok = row < rows;
    float v[NQ][4];
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      uint2 raw = make_uint2(0u, 0u);
      if (ok) raw = *reinterpret_cast<const uint2*>(x + row * C + 4 * (lg + G * q));
      unpack4(raw, v[q]);
      s += v[q][0] + v[q][1] + v[q][2] + v[q][3];
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.f / C);
    float qq = 0.f;
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[q][i] -= mean;
        qq += v[q][i] * v[q][i];
      }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) qq += __shfl_xor_sync(0xffffffffu, qq, o);
    const float rstd = rsqrtf(qq * (1.f / C) + eps);
    if (ok) {
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        uint2 o2;
        o2.x = pack2(v[q][0] * rstd * gm[q][0] + bt[q][0], v[q][1] * rstd * gm[q][1] + bt[q][1]);
        o2.y = pack2(v[q][2] * rstd * gm[q][2] + bt[q][2], v[q][3] * rstd * gm[q][3] + bt[q][3]);
        *reinterpret_cast<uint2*>(y + row * C + 4 * (lg + G * q)) = o2;
      }
      if (lg == 0 && mean_out) {
        mean_out[row] = mean;
        rstd_out[row] = rstd;
      }
    }
  }
}

template <typename TI>
__global__ void __launch_bounds__(256) im2col_rows_kernel(const TI* __restrict__ x, bf16* __restrict__ cols, int Cin, int T,
                                                          int H, int W, int To, int Ho, int Wo, int kt, int kh, int kw,
                                                          int st, int sh, int sw, int pt, int ph, int pw, int Kpad, int RW) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int R = Cin * kt * kh;
  bf16* rows = reinterpret_cast<bf16*>(smem_raw);                    // [R][RW], halo of pw zeros on both sides
  int* koff = reinterpret_cast<int*>(smem_raw + (((size_t)R * RW * 2 + 15) & ~size_t(15)));  // [Kpad]
  const int ho = blockIdx.x % Ho, to = (blockIdx.x / Ho) % To, b = blockIdx.x / (Ho * To);
  const int K = R * kw;
  for (int k = threadIdx.x; k < Kpad; k += blockDim.x) koff[k] = k < K ? (k / kw) * RW + (k % kw) : -1;
  for (int i = threadIdx.x; i < R * RW; i += blockDim.x) {
    const int r = i / RW, col = i % RW;
    const int dh = r % kh, dt = (r / kh) % kt, c = r / (kh * kt);
    const int t = to * st - pt + dt, hh = ho * sh - ph + dh, ww = col - pw;
    float v = 0.f;
    if (t >= 0 && t < T && hh >= 0 && hh < H && ww >= 0 && ww < W)
      v = to_f(x[((((int64_t)b * Cin + c) * T + t) * H + hh) * W + ww]);
    rows[i] = __float2bfloat16_rn(v);
  }
  __syncthreads();
  const int cpr = Kpad >> 3;  // 16-byte chunks per output row
  bf16* obase = cols + ((((int64_t)b * To + to) * Ho + ho) * Wo) * Kpad;
  const unsigned short* rs = reinterpret_cast<const unsigned short*>(rows);
  for (int i = threadIdx.x; i < Wo * cpr; i += blockDim.x) {
    const int wo = i / cpr, q = i % cpr;
    const int base = wo * sw;
    unsigned short e[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int off = koff[q * 8 + u];
      e[u] = off < 0 ? (unsigned short)0 : rs[off + base];
    }
    uint4 o;
    o.x = e[0] | ((uint32_t)e[1] << 16); o.y = e[2] | ((uint32_t)e[3] << 16);
    o.z = e[4] | ((uint32_t)e[5] << 16); o.w = e[6] | ((uint32_t)e[7] << 16);
    *reinterpret_cast<uint4*>(obase + (int64_t)wo * Kpad + q * 8) = o;
  }
}

// Fast variant for bf16 clips: CTA = 4 * (Kpad / 8) threads per (b, t', h') output row.  The R = Cin*kt*kh input
// rows are staged with 16-byte loads (halo of HALO zero elements on both sides); a thread owns ONE 16-byte column
// chunk q of the lowered rows, so the eight source offsets of its chunk sit in registers and every output row
// (Kpad * 2 bytes) is written as contiguous 16-byte stores.
constexpr int IM_HALO = 4;
__global__ void __launch_bounds__(256) im2col_rows_vec_kernel(const bf16* __restrict__ x, bf16* __restrict__ cols, int Cin,
                                                              int T, int H, int W, int To, int Ho, int Wo, int kt, int kh,
                                                              int kw, int st, int sh, int sw, int pt, int ph, int pw,
                                                              int Kpad, int RW) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned short* rows = reinterpret_cast<unsigned short*>(smem_raw);  // [R][RW]
  const int R = Cin * kt * kh, K = R * kw;
  const int ho = blockIdx.x % Ho, to = (blockIdx.x / Ho) % To, b = blockIdx.x / (Ho * To);
  const int cpr = Kpad >> 3, wchunks = W >> 3;
  // stage: row r = (c, dt, dh); 16-byte chunks of the W input pixels land at element HALO + 8*j (8-byte aligned)
  for (int i = threadIdx.x; i < R * (wchunks + 2); i += blockDim.x) {
    const int r = i / (wchunks + 2), j = i % (wchunks + 2) - 1;
    unsigned short* dst = rows + r * RW;
    if (j < 0) {
      *reinterpret_cast<uint2*>(dst) = make_uint2(0u, 0u);
    } else if (j == wchunks) {
      *reinterpret_cast<uint2*>(dst + IM_HALO + W) = make_uint2(0u, 0u);
    } else {
      const int dh = r % kh, dt = (r / kh) % kt, c = r / (kh * kt);
      const int t = to * st - pt + dt, hh = ho * sh - ph + dh;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (t >= 0 && t < T && hh >= 0 && hh < H)
        v = __ldg(reinterpret_cast<const uint4*>(x + ((((int64_t)b * Cin + c) * T + t) * H + hh) * W + 8 * j));
      uint2* d2 = reinterpret_cast<uint2*>(dst + IM_HALO + 8 * j);
      d2[0] = make_uint2(v.x, v.y);
      d2[1] = make_uint2(v.z, v.w);
    }
  }
  const int q = threadIdx.x % cpr, w_first = threadIdx.x / cpr, w_step = blockDim.x / cpr;
  int off[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int k = q * 8 + u;
    off[u] = k < K ? (k / kw) * RW + (k % kw) + IM_HALO - pw : -1;
  }
  __syncthreads();
  bf16* obase = cols + ((((int64_t)b * To + to) * Ho + ho) * Wo) * Kpad + q * 8;
  for (int wo = w_first; wo < Wo; wo += w_step) {
    const int base = wo * sw;
    unsigned short e[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) e[u] = off[u] < 0 ? (unsigned short)0 : rows[off[u] + base];
    uint4 o;
    o.x = e[0] | ((uint32_t)e[1] << 16); o.y = e[2] | ((uint32_t)e[3] << 16);
    o.z = e[4] | ((uint32_t)e[5] << 16); o.w = e[6] | ((uint32_t)e[7] << 16);
    *reinterpret_cast<uint4*>(obase + (int64_t)wo * Kpad) = o;
  }
}

// skip-path MaxPool3d k(1,3,3) s(1,s,s) p(0,1,1) on [B, N, C] (attention.py:562-564): one thread per 8 channels
// IDX: also record, per output element, which of the 9 window positions (dh * 3 + dw, ATen's scan order, first maximum)
// won -- the backward then compares one byte instead of re-scanning up to four windows of nine values per input element.
template <bool IDX>
__global__ void __launch_bounds__(256) skip_maxpool_bf16_kernel(const bf16* __restrict__ x, bf16* __restrict__ y,
                                                                uint8_t* __restrict__ idx, int B, int C,
                                                                int T_, int H, int W, int Ho, int Wo, int O, int s) {
  const int C8 = C >> 3;
  const int64_t Lin = (int64_t)T_ * H * W, Lout = (int64_t)T_ * Ho * Wo;
  const int64_t Nin = 1 + Lin + O, Nout = 1 + Lout + O;
  const int64_t total = (int64_t)B * Nout * C8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    const int64_t r = i / C8;
    const int64_t tok = r % Nout;
    const int b = (int)(r / Nout);
    const bf16* xb = x + (int64_t)b * Nin * C + c8 * 8;
    uint4 o;
    if (tok == 0) {
      o = *reinterpret_cast<const uint4*>(xb);
    } else if (tok > Lout) {
      o = *reinterpret_cast<const uint4*>(xb + (tok - Lout + Lin) * C);
    } else {
      const int64_t p = tok - 1;
      const int wo = (int)(p % Wo), ho = (int)((p / Wo) % Ho), t = (int)(p / ((int64_t)Wo * Ho));
      uint4 taps[9];
      bool ok[9];
#pragma unroll
      for (int dh = 0; dh < 3; ++dh)
#pragma unroll
        for (int dw = 0; dw < 3; ++dw) {
          const int hh = ho * s - 1 + dh, ww = wo * s - 1 + dw;
          ok[dh * 3 + dw] = hh >= 0 && hh < H && ww >= 0 && ww < W;
          taps[dh * 3 + dw] = ok[dh * 3 + dw]
              ? *reinterpret_cast<const uint4*>(xb + (1 + ((int64_t)t * H + hh) * W + ww) * C) : make_uint4(0, 0, 0, 0);
        }
      float m[8];
      uint32_t win[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { m[u] = -INFINITY; win[u] = 0u; }
#pragma unroll
      for (int k = 0; k < 9; ++k)
        if (ok[k]) {
          const uint32_t wv[4] = {taps[k].x, taps[k].y, taps[k].z, taps[k].w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float a = __uint_as_float(wv[u] << 16), c = __uint_as_float(wv[u] & 0xffff0000u);
            if (a > m[2 * u] || a != a) { m[2 * u] = a; if (IDX) win[2 * u] = (uint32_t)k; }
            if (c > m[2 * u + 1] || c != c) { m[2 * u + 1] = c; if (IDX) win[2 * u + 1] = (uint32_t)k; }
          }
        }
      o.x = pack2(m[0], m[1]); o.y = pack2(m[2], m[3]); o.z = pack2(m[4], m[5]); o.w = pack2(m[6], m[7]);
      if (IDX) {
        uint2 iv;
        iv.x = win[0] | (win[1] << 8) | (win[2] << 16) | (win[3] << 24);
        iv.y = win[4] | (win[5] << 8) | (win[6] << 16) | (win[7] << 24);
        *reinterpret_cast<uint2*>(idx + ((int64_t)b * Nout + tok) * C + c8 * 8) = iv;
      }
    }
    *reinterpret_cast<uint4*>(y + ((int64_t)b * Nout + tok) * C + c8 * 8) = o;
  }
}

// Backward with the recorded window positions: thread = 8 channels of one INPUT token; for each of the (at most four)
// windows that contain the token it compares the eight recorded positions with its own and takes dy where they match.
__global__ void __launch_bounds__(256) skip_maxpool_bwd_idx_kernel(const uint8_t* __restrict__ idx, const bf16* __restrict__ dy,
                                                                   bf16* __restrict__ dx, int B, int C, int T_, int H, int W,
                                                                   int Ho, int Wo, int O, int s) {
  const int C8 = C >> 3;
  const int64_t Lin = (int64_t)T_ * H * W, Lout = (int64_t)T_ * Ho * Wo;
  const int64_t Nin = 1 + Lin + O, Nout = 1 + Lout + O;
  const int64_t total = (int64_t)B * Nin * C8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8) * 8;
    const int64_t r = i / C8;
    const int64_t tok = r % Nin;
    const int b = (int)(r / Nin);
    const bf16* dyb = dy + (int64_t)b * Nout * C + c;
    const uint8_t* ib = idx + (int64_t)b * Nout * C + c;
    bf16* o = dx + r * C + c;
    if (tok == 0) {
      *reinterpret_cast<uint4*>(o) = __ldg(reinterpret_cast<const uint4*>(dyb));
      continue;
    }
    if (tok > Lin) {
      *reinterpret_cast<uint4*>(o) = __ldg(reinterpret_cast<const uint4*>(dyb + (tok - Lin + Lout) * C));
      continue;
    }
    const int64_t p = tok - 1;
    const int w = (int)(p % W), h = (int)((p / W) % H), t = (int)(p / ((int64_t)W * H));
    float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int ho = (h - 1 + s - 1) / s; ho * s <= h + 1 && ho < Ho; ++ho) {
      if (ho < 0) continue;
      const int dh = h - (ho * s - 1);
      for (int wo = (w - 1 + s - 1) / s; wo * s <= w + 1 && wo < Wo; ++wo) {
        if (wo < 0) continue;
        const uint32_t me = (uint32_t)(dh * 3 + (w - (wo * s - 1)));
        const int64_t orow = 1 + ((int64_t)t * Ho + ho) * Wo + wo;
        const uint2 iv = __ldg(reinterpret_cast<const uint2*>(ib + orow * C));
        const uint32_t hit_lo = iv.x ^ (me * 0x01010101u), hit_hi = iv.y ^ (me * 0x01010101u);  // zero byte = match
        if (((hit_lo - 0x01010101u) & ~hit_lo & 0x80808080u) | ((hit_hi - 0x01010101u) & ~hit_hi & 0x80808080u)) {
          const uint4 dv = __ldg(reinterpret_cast<const uint4*>(dyb + orow * C));
          const uint32_t wv[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t b0 = ((u < 2 ? hit_lo : hit_hi) >> (16 * (u & 1))) & 0xffu;
            const uint32_t b1 = ((u < 2 ? hit_lo : hit_hi) >> (16 * (u & 1) + 8)) & 0xffu;
            if (b0 == 0u) g[2 * u] += __uint_as_float(wv[u] << 16);
            if (b1 == 0u) g[2 * u + 1] += __uint_as_float(wv[u] & 0xffff0000u);
          }
        }
      }
    }
    uint4 ov;
    ov.x = pack2(g[0], g[1]); ov.y = pack2(g[2], g[3]); ov.z = pack2(g[4], g[5]); ov.w = pack2(g[6], g[7]);
    *reinterpret_cast<uint4*>(o) = ov;
  }
}

}  // namespace

extern "C" int svit_skip_maxpool_fwd_idx(const void* x, void* y, void* idx, int B, int C, int T, int H, int W, int O,
                                         int stride_hw, int dtype, void* stream) {
  if (!x || !y || !idx || stride_hw < 2 || B < 0 || O < 1) return SVIT_EINVAL;
  if (dtype != SVIT_BF16 || C % 8 ||
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) || (reinterpret_cast<uintptr_t>(idx) & 7))
    return SVIT_ENOTSUP;
  const int Ho = (H - 1) / stride_hw + 1, Wo = (W - 1) / stride_hw + 1;
  const int64_t total = (int64_t)B * (1 + (int64_t)T * Ho * Wo + O) * (C / 8);
  if (total == 0) return 0;
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)svit_num_sms() * 32) blocks = (int64_t)svit_num_sms() * 32;
  skip_maxpool_bf16_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, (uint8_t*)idx, B, C,
                                                                                   T, H, W, Ho, Wo, O, stride_hw);
  SVIT_CHECK_LAUNCH();
  return 0;
}

extern "C" int svit_skip_maxpool_bwd_idx(const void* idx, const void* dy, void* dx, int B, int C, int T, int H, int W, int O,
                                         int stride_hw, int dtype, void* stream) {
  if (!idx || !dy || !dx || stride_hw < 2 || B < 0 || O < 1) return SVIT_EINVAL;
  if (dtype != SVIT_BF16 || C % 8 ||
      ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) || (reinterpret_cast<uintptr_t>(idx) & 7))
    return SVIT_ENOTSUP;
  const int Ho = (H - 1) / stride_hw + 1, Wo = (W - 1) / stride_hw + 1;
  const int64_t total = (int64_t)B * (1 + (int64_t)T * H * W + O) * (C / 8);
  if (total == 0) return 0;
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)svit_num_sms() * 32) blocks = (int64_t)svit_num_sms() * 32;
  skip_maxpool_bwd_idx_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)idx, (const bf16*)dy, (bf16*)dx, B,
                                                                                C, T, H, W, Ho, Wo, O, stride_hw);
  SVIT_CHECK_LAUNCH();
  return 0;
}

int svit_skip_maxpool_fwd_bf16(const void* x, void* y, int B, int C, int T, int H, int W, int Ho, int Wo, int O, int s,
                               cudaStream_t st) {
  if (C % 8 || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15)) return 0;
  const int64_t total = (int64_t)B * (1 + (int64_t)T * Ho * Wo + O) * (C / 8);
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)svit_num_sms() * 32) blocks = (int64_t)svit_num_sms() * 32;
  skip_maxpool_bf16_kernel<false><<<(unsigned)blocks, 256, 0, st>>>((const bf16*)x, (bf16*)y, nullptr, B, C, T, H, W, Ho, Wo, O, s);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : 1000 + (int)e;
}

// returns 1 if handled (launched), 0 if the shape is not covered by the fast path, > 1 on CUDA error (offset by 1000)
int svit_layernorm_fwd_bf16(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                            int64_t rows, int C, float eps, cudaStream_t st) {
  if ((reinterpret_cast<uintptr_t>(x) & 7) || (reinterpret_cast<uintptr_t>(y) & 7) ||
      (reinterpret_cast<uintptr_t>(gamma) & 15) || (reinterpret_cast<uintptr_t>(beta) & 15))
    return 0;
  const int sms = svit_num_sms();
#define LN_LAUNCH(G, NQ)                                                                                           \
  {                                                                                                                \
    int64_t blocks = (rows * G + 255) / 256;                                                                       \
    if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;                                                    \
    layernorm_bf16_kernel<G, NQ><<<(unsigned)blocks, 256, 0, st>>>((const bf16*)x, gamma, beta, (bf16*)y, mean, rstd, \
                                                                  rows, eps);                                      \
  }
  if (C == 96) LN_LAUNCH(8, 3)
  else if (C == 192) LN_LAUNCH(16, 3)
  else if (C == 384) LN_LAUNCH(32, 3)
  else if (C == 768) LN_LAUNCH(32, 6)
  else return 0;
#undef LN_LAUNCH
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : 1000 + (int)e;
}

int svit_im2col_rows(const void* x, void* cols, int B, int Cin, int T, int H, int W, int To, int Ho, int Wo, int kt, int kh,
                     int kw, int st_, int sh, int sw, int pt, int ph, int pw, int Kpad, int in_dtype, cudaStream_t st) {
  if (Kpad % 8 || (reinterpret_cast<uintptr_t>(cols) & 15)) return 0;
  const int R = Cin * kt * kh;
  const int RW = (W + 2 * pw + 1) & ~1;
  if ((Wo - 1) * sw + kw > RW) return 0;
  const size_t smem = (((size_t)R * RW * 2 + 15) & ~size_t(15)) + (size_t)Kpad * 4;
  if (smem > 200 * 1024) return 0;
  const unsigned grid = (unsigned)((int64_t)B * To * Ho);
  cudaError_t e;
  const int cpr = Kpad / 8;
  if (in_dtype == SVIT_BF16 && W % 8 == 0 && pw <= IM_HALO && cpr <= 64 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const int RWv = (W + 2 * IM_HALO + 7) & ~7;  // rows 16-byte aligned
    const int threads = (256 / cpr) * cpr;
    const size_t smv = (size_t)R * RWv * 2;
    if ((Wo - 1) * sw + kw - 1 - pw < W + 4 && smv <= 200 * 1024) {
      static size_t confv = 0;
      if (smv > confv) {
        if ((e = cudaFuncSetAttribute(im2col_rows_vec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smv)) != cudaSuccess)
          return 1000 + (int)e;
        confv = smv;
      }
      im2col_rows_vec_kernel<<<grid, threads, smv, st>>>((const bf16*)x, (bf16*)cols, Cin, T, H, W, To, Ho, Wo, kt, kh, kw, st_,
                                                         sh, sw, pt, ph, pw, Kpad, RWv);
      e = cudaGetLastError();
      return e == cudaSuccess ? 1 : 1000 + (int)e;
    }
  }
  if (in_dtype == SVIT_BF16) {
    static size_t conf = 0;
    if (smem > conf) {
      if ((e = cudaFuncSetAttribute(im2col_rows_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
        return 1000 + (int)e;
      conf = smem;
    }
    im2col_rows_kernel<bf16><<<grid, 256, smem, st>>>((const bf16*)x, (bf16*)cols, Cin, T, H, W, To, Ho, Wo, kt, kh, kw, st_,
                                                      sh, sw, pt, ph, pw, Kpad, RW);
  } else {
    static size_t conf = 0;
    if (smem > conf) {
      if ((e = cudaFuncSetAttribute(im2col_rows_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
        return 1000 + (int)e;
      conf = smem;
    }
    im2col_rows_kernel<float><<<grid, 256, smem, st>>>((const float*)x, (bf16*)cols, Cin, T, H, W, To, Ho, Wo, kt, kh, kw,
                                                       st_, sh, sw, pt, ph, pw, Kpad, RW);
  }
  e = cudaGetLastError();
  return e == cudaSuccess ? 1 : 1000 + (int)e;
}


// ------------------------------------------------------------------------------------------------ uint8 frames
// Input side of the path: decoded frames arrive as uint8 [B, T, H, W, 3]; the reference normalises on
// the CPU (datasets/utils.py:287-303: x / 255, - mean, / std in fp32) and permutes to [B, 3, T, H, W] before the
// host -> device copy of 4 bytes per value.  Here the uint8 frames are what crosses PCIe (1 byte per value) and one
// streaming kernel does normalise + layout change + cast: thread = 16 consecutive pixels of a row (48 bytes in, 3 x 16
// values out), same fp32 operation order as the reference so the fp32 output is bit-identical.
template <typename OT>
__global__ void __launch_bounds__(256) normalize_u8_kernel(const uint8_t* __restrict__ in, OT* __restrict__ out, int64_t npix16,
                                                           int64_t THW, float m0, float m1, float m2, float s0, float s1,
                                                           float s2) {
  const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix16; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i * 16;          // first pixel (b, t, h, w) flattened; THW % 16 == 0 keeps a unit inside one sample
    const int64_t b = pix / THW, r = pix - b * THW;
    const uint4* src = reinterpret_cast<const uint4*>(in + pix * 3);
    const uint4 q0 = __ldg(src), q1 = __ldg(src + 1), q2 = __ldg(src + 2);
    const uint32_t w[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
    float v[3][16];
#pragma unroll
    for (int k = 0; k < 48; ++k) {
      const float f = (float)((w[k >> 2] >> ((k & 3) * 8)) & 255u);
      const int c = k % 3;
      v[c][k / 3] = __fdiv_rn(__fsub_rn(__fdiv_rn(f, 255.0f), mean[c]), sd[c]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      OT* dst = out + (b * 3 + c) * THW + r;
      if (sizeof(OT) == 2) {
        uint32_t o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          __nv_bfloat162 h = __floats2bfloat162_rn(v[c][2 * k], v[c][2 * k + 1]);
          o[k] = *reinterpret_cast<uint32_t*>(&h);
        }
        reinterpret_cast<uint4*>(dst)[0] = make_uint4(o[0], o[1], o[2], o[3]);
        reinterpret_cast<uint4*>(dst)[1] = make_uint4(o[4], o[5], o[6], o[7]);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          reinterpret_cast<float4*>(dst)[k] = make_float4(v[c][4 * k], v[c][4 * k + 1], v[c][4 * k + 2], v[c][4 * k + 3]);
      }
    }
  }
}

extern "C" int svit_normalize_u8(const void* frames, void* out, int B, int T, int H, int W, float mean0, float mean1,
                                 float mean2, float std0, float std1, float std2, int out_dtype, void* stream) {
  if (!frames || !out || B < 0 || T < 1 || H < 1 || W < 1) return SVIT_EINVAL;
  const int64_t THW = (int64_t)T * H * W;
  if (THW % 16 || (reinterpret_cast<uintptr_t>(frames) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return SVIT_ENOTSUP;
  if (B == 0) return 0;
  const int64_t n16 = (int64_t)B * THW / 16;
  cudaStream_t st = (cudaStream_t)stream;
  int64_t grid = (n16 + 255) / 256;
  const int64_t cap = (int64_t)svit_num_sms() * 16;
  if (grid > cap) grid = cap;
  if (out_dtype == SVIT_BF16)
    normalize_u8_kernel<bf16><<<(unsigned)grid, 256, 0, st>>>((const uint8_t*)frames, (bf16*)out, n16, THW, mean0, mean1, mean2, std0, std1, std2);
  else if (out_dtype == SVIT_F32)
    normalize_u8_kernel<float><<<(unsigned)grid, 256, 0, st>>>((const uint8_t*)frames, (float*)out, n16, THW, mean0, mean1, mean2, std0, std1, std2);
  else
    return SVIT_EINVAL;
  SVIT_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------ N4: crop / flip
// Input side on the GPU: decoded uint8 frames [B, T, H, W, 3] -> per-sample spatial crop
// (datasets/transform.py:154-190 random_crop / uniform_crop with the offsets chosen by the host), optional horizontal
// flip (transform.py:248-285), colour normalisation (datasets/utils.py:287-303) and the THWC -> CTHW layout, one pass.
// thread = 4 consecutive output pixels of a row; fp32 operation order of the reference (bit-identical fp32 output).
template <typename OT>
__global__ void __launch_bounds__(256) crop_flip_normalize_u8_kernel(const uint8_t* __restrict__ in, OT* __restrict__ out,
                                                                     const int32_t* __restrict__ x_off,
                                                                     const int32_t* __restrict__ y_off,
                                                                     const int32_t* __restrict__ flip, int64_t nquads, int T,
                                                                     int H, int W, int cs_h, int cs_w, float m0, float m1,
                                                                     float m2, float s0, float s1, float s2) {
  const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
  const int qw = cs_w >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nquads; i += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(i % qw);
    int64_t r = i / qw;
    const int y = (int)(r % cs_h);
    r /= cs_h;
    const int t = (int)(r % T);
    const int b = (int)(r / T);
    const int xo = x_off[b], yo = y_off[b];
    const bool fl = flip && flip[b] != 0;
    const uint8_t* row = in + ((((int64_t)b * T + t) * H + (y + yo)) * W) * 3;
    float v[3][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int x = q * 4 + k;
      const int sx = xo + (fl ? cs_w - 1 - x : x);  // images.flip(-1) after the crop
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float f = (float)__ldg(row + (int64_t)sx * 3 + c);
        v[c][k] = __fdiv_rn(__fsub_rn(__fdiv_rn(f, 255.0f), mean[c]), sd[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      OT* dst = out + ((((int64_t)b * 3 + c) * T + t) * cs_h + y) * cs_w + q * 4;
      if (sizeof(OT) == 2) {
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[c][0], v[c][1]), h1 = __floats2bfloat162_rn(v[c][2], v[c][3]);
        *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
      } else {
        *reinterpret_cast<float4*>(dst) = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
      }
    }
  }
}

// Boxes follow the frames (transform.py:107-132 crop_clip_boxes, :248-285 horizontal_flip on [N, 4] boxes) and are then
// brought to the loss's format (datasets/ssv2_frames.py:347-353): normalise by the crop size, clip to [0, 1],
// xyxy -> cxcywh (utils/box_ops.py:32-36), zero boxes with w or h <= eps (utils/box_ops.py:116-130).  Every step is the
// reference's float32 operation, so the result is bit-identical.
__global__ void boxes_crop_flip_kernel(const float* __restrict__ in, float* __restrict__ out, const int32_t* __restrict__ x_off,
                                       const int32_t* __restrict__ y_off, const int32_t* __restrict__ flip, int64_t n,
                                       int per_sample, int cs_h, int cs_w, float eps) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / per_sample);
    const float xo = (float)x_off[b], yo = (float)y_off[b];
    const float W = (float)cs_w, Hh = (float)cs_h;
    float x0 = fminf(fmaxf(__fsub_rn(in[4 * i + 0], xo), 0.f), W), y0 = fminf(fmaxf(__fsub_rn(in[4 * i + 1], yo), 0.f), Hh);
    float x1 = fminf(fmaxf(__fsub_rn(in[4 * i + 2], xo), 0.f), W), y1 = fminf(fmaxf(__fsub_rn(in[4 * i + 3], yo), 0.f), Hh);
    if (flip && flip[b] != 0) {
      const float nx0 = __fsub_rn(__fsub_rn(W, x1), 1.f), nx1 = __fsub_rn(__fsub_rn(W, x0), 1.f);
      x0 = nx0;
      x1 = nx1;
    }
    x0 = fminf(fmaxf(__fdiv_rn(x0, W), 0.f), 1.f);
    x1 = fminf(fmaxf(__fdiv_rn(x1, W), 0.f), 1.f);
    y0 = fminf(fmaxf(__fdiv_rn(y0, Hh), 0.f), 1.f);
    y1 = fminf(fmaxf(__fdiv_rn(y1, Hh), 0.f), 1.f);
    float cx = __fdiv_rn(__fadd_rn(x0, x1), 2.f), cy = __fdiv_rn(__fadd_rn(y0, y1), 2.f);
    float w = __fsub_rn(x1, x0), h = __fsub_rn(y1, y0);
    if (w <= eps || h <= eps) cx = cy = w = h = 0.f;
    out[4 * i + 0] = cx;
    out[4 * i + 1] = cy;
    out[4 * i + 2] = w;
    out[4 * i + 3] = h;
  }
}

extern "C" int svit_crop_flip_normalize_u8(const void* frames, void* out, const int32_t* x_off, const int32_t* y_off,
                                           const int32_t* flip, int B, int T, int H, int W, int crop_h, int crop_w,
                                           float mean0, float mean1, float mean2, float std0, float std1, float std2,
                                           int out_dtype, void* stream) {
  if (!frames || !out || !x_off || !y_off || B < 0 || T < 1 || H < 1 || W < 1 || crop_h < 1 || crop_w < 1 || crop_h > H ||
      crop_w > W)
    return SVIT_EINVAL;
  if (crop_w % 4 || (reinterpret_cast<uintptr_t>(out) & 15)) return SVIT_ENOTSUP;
  if (B == 0) return 0;
  const int64_t nq = (int64_t)B * T * crop_h * (crop_w / 4);
  cudaStream_t st = (cudaStream_t)stream;
  int64_t grid = (nq + 255) / 256;
  const int64_t cap = (int64_t)svit_num_sms() * 16;
  if (grid > cap) grid = cap;
  if (out_dtype == SVIT_BF16)
    crop_flip_normalize_u8_kernel<bf16><<<(unsigned)grid, 256, 0, st>>>((const uint8_t*)frames, (bf16*)out, x_off, y_off, flip, nq, T, H, W, crop_h, crop_w, mean0, mean1, mean2, std0, std1, std2);
  else if (out_dtype == SVIT_F32)
    crop_flip_normalize_u8_kernel<float><<<(unsigned)grid, 256, 0, st>>>((const uint8_t*)frames, (float*)out, x_off, y_off, flip, nq, T, H, W, crop_h, crop_w, mean0, mean1, mean2, std0, std1, std2);
  else
    return SVIT_EINVAL;
  SVIT_CHECK_LAUNCH();
  return 0;
}

extern "C" int svit_boxes_crop_flip(const float* boxes_xyxy, float* out_cxcywh, const int32_t* x_off, const int32_t* y_off,
                                    const int32_t* flip, int B, int64_t boxes_per_sample, int crop_h, int crop_w, float eps,
                                    void* stream) {
  if (!boxes_xyxy || !out_cxcywh || !x_off || !y_off || B < 0 || boxes_per_sample < 0 || crop_h < 1 || crop_w < 1)
    return SVIT_EINVAL;
  const int64_t n = (int64_t)B * boxes_per_sample;
  if (n == 0) return 0;
  if (boxes_per_sample > 0x7fffffff) return SVIT_EINVAL;
  boxes_crop_flip_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(boxes_xyxy, out_cxcywh, x_off, y_off, flip, n, (int)boxes_per_sample, crop_h, crop_w, eps);
  SVIT_CHECK_LAUNCH();
  return 0;
}


// ------------------------------------------------------------------------------------------------ row statistics
// (mean, rstd) per row for a LayerNorm folded into the consuming GEMM (gemm_tc2.cu).  Same work split as
// layernorm_bf16_kernel (G lanes per row, NQ bf16x4 quads per lane, the row stays in registers between the mean pass and
// the centred second-moment pass) without the write-back: half the traffic of a LayerNorm pass.
template <int G, int NQ>
__global__ void __launch_bounds__(256) row_stats_bf16_kernel(const bf16* __restrict__ x, float2* __restrict__ stats,
                                                             int64_t rows, float eps) {
  constexpr int C = 4 * G * NQ;
  const int lg = threadIdx.x % G;
  const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  const int64_t ngroups = (int64_t)gridDim.x * blockDim.x / G;
  const int64_t iters = (rows + ngroups - 1) / ngroups;
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t row = group + it * ngroups;
    const bool ok = row < rows;
    float v[NQ][4];
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      uint2 raw = make_uint2(0u, 0u);
      if (ok) raw = __ldg(reinterpret_cast<const uint2*>(x + row * C + 4 * (lg + G * q)));
      unpack4(raw, v[q]);
      s += v[q][0] + v[q][1] + v[q][2] + v[q][3];
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.f / C);
    float qq = 0.f;
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float d = v[q][i] - mean;
        qq = fmaf(d, d, qq);
      }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) qq += __shfl_xor_sync(0xffffffffu, qq, o);
    if (ok && lg == 0) stats[row] = make_float2(mean, rsqrtf(qq * (1.f / C) + eps));
  }
}

// generic widths / fp32 rows: warp per row, 16-byte loads (C <= 128 vectors)
template <typename T>
__global__ void __launch_bounds__(256) row_stats_kernel(const T* __restrict__ x, float2* __restrict__ stats, int64_t M, int C,
                                                        float eps) {
  constexpr int V = 16 / sizeof(T);  // elements per 16-byte load
  const int lane = threadIdx.x & 31;
  const int nv = C / V;              // vectors per row
  for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < M; row += (int64_t)gridDim.x * 8) {
    const uint4* src = reinterpret_cast<const uint4*>(x + row * C);
    float v[4][8];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = lane + 32 * k;
      if (i < nv) {
        const uint4 q = __ldg(src + i);
        if (sizeof(T) == 2) {
          const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            v[k][2 * u] = __uint_as_float(w[u] << 16);
            v[k][2 * u + 1] = __uint_as_float(w[u] & 0xffff0000u);
          }
        } else {
          v[k][0] = __uint_as_float(q.x); v[k][1] = __uint_as_float(q.y);
          v[k][2] = __uint_as_float(q.z); v[k][3] = __uint_as_float(q.w);
        }
#pragma unroll
        for (int u = 0; u < V; ++u) s += v[k][u];
      }
    }
    const float mean = warp_sum(s) / (float)C;
    float qsum = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (lane + 32 * k < nv) {
#pragma unroll
        for (int u = 0; u < V; ++u) {
          const float d = v[k][u] - mean;
          qsum = fmaf(d, d, qsum);
        }
      }
    }
    const float var = warp_sum(qsum) / (float)C;
    if (lane == 0) stats[row] = make_float2(mean, rsqrtf(var + eps));
  }
}

extern "C" int svit_row_stats(const void* x, float* stats, int64_t M, int C, float eps, int dtype, void* stream) {
  if (!x || !stats || M < 0 || C < 1) return SVIT_EINVAL;
  if (M == 0) return 0;
  const int V = dtype == SVIT_BF16 ? 8 : 4;
  if ((dtype != SVIT_BF16 && dtype != SVIT_F32) || C % V || C / V > 128 || (reinterpret_cast<uintptr_t>(x) & 15) ||
      (reinterpret_cast<uintptr_t>(stats) & 7))
    return SVIT_ENOTSUP;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t cap = (int64_t)svit_num_sms() * 16;
#define RS_LAUNCH(G, NQ)                                                                                        \
  {                                                                                                             \
    int64_t blocks = (M * G + 255) / 256;                                                                       \
    if (blocks > cap) blocks = cap;                                                                             \
    row_stats_bf16_kernel<G, NQ><<<(unsigned)blocks, 256, 0, st>>>((const bf16*)x, (float2*)stats, M, eps);     \
  }
  if (dtype == SVIT_BF16 && (C == 96 || C == 192 || C == 384 || C == 768)) {
    if (C == 96) RS_LAUNCH(8, 3)
    else if (C == 192) RS_LAUNCH(16, 3)
    else if (C == 384) RS_LAUNCH(32, 3)
    else RS_LAUNCH(32, 6)
  } else {
    int64_t grid = (M + 7) / 8;
    if (grid > cap) grid = cap;
    if (dtype == SVIT_BF16) row_stats_kernel<bf16><<<(unsigned)grid, 256, 0, st>>>((const bf16*)x, (float2*)stats, M, C, eps);
    else row_stats_kernel<float><<<(unsigned)grid, 256, 0, st>>>((const float*)x, (float2*)stats, M, C, eps);
  }
#undef RS_LAUNCH
  SVIT_CHECK_LAUNCH();
  return 0;
}
