// Box-conditioned object tokens: per-frame RoIAlign over the patch grid + max over bins, and the
// integer box -> slot rules.  Reference call sites: video_model_builder.py:385-392, 472-491 (RoIAlign
// 7x7, scale 1/16, aligned=True, adaptive sampling); utils/box_ops.py:116-130, 140-194.
// RoIAlign arithmetic follows torchvision.ops.roi_align (the reference's head_helper.py is absent).
// Gather-bound: one CTA per box, threads over channels so every bilinear tap is a coalesced row read.
#include "common.cuh"

__device__ __forceinline__ void bilinear_setup(float y, float x, int H, int W, int& yl, int& xl, int& yh, int& xh,
                                               float& w1, float& w2, float& w3, float& w4, bool& valid) {
  valid = !(y < -1.0f || y > (float)H || x < -1.0f || x > (float)W);
  if (!valid) return;
  if (y <= 0.f) y = 0.f;
  if (x <= 0.f) x = 0.f;
  yl = (int)y;
  xl = (int)x;
  if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else { yh = yl + 1; }
  if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else { xh = xl + 1; }
  float ly = y - yl, lx = x - xl, hy = 1.f - ly, hx = 1.f - lx;
  w1 = hy * hx; w2 = hy * lx; w3 = ly * hx; w4 = ly * lx;
}

// value of one RoIAlign bin for channel c; fm points at pixel (0,0) channel 0 of a channels-last [H, W, C] map
template <typename T>
__device__ __forceinline__ float roi_bin(const T* __restrict__ fm, int C, int c, int H, int W, float y1, float x1,
                                         float bh, float bw, int gh, int gw, int ph, int pw) {
  float acc = 0.f;
  for (int iy = 0; iy < gh; ++iy) {
    float y = y1 + ph * bh + (iy + 0.5f) * bh / (float)gh;
    for (int ix = 0; ix < gw; ++ix) {
      float x = x1 + pw * bw + (ix + 0.5f) * bw / (float)gw;
      int yl, xl, yh, xh;
      float w1, w2, w3, w4;
      bool valid;
      bilinear_setup(y, x, H, W, yl, xl, yh, xh, w1, w2, w3, w4, valid);
      if (!valid) continue;
      acc += w1 * to_f(fm[((int64_t)yl * W + xl) * C + c]) + w2 * to_f(fm[((int64_t)yl * W + xh) * C + c]) +
             w3 * to_f(fm[((int64_t)yh * W + xl) * C + c]) + w4 * to_f(fm[((int64_t)yh * W + xh) * C + c]);
    }
  }
  int cnt = gh * gw;
  return acc / (float)(cnt > 0 ? cnt : 1);
}

template <typename T>
__global__ void roi_tokens_kernel(const T* __restrict__ feat, int64_t feat_bs, const float* __restrict__ boxes,
                                  T* __restrict__ tokens, int64_t tok_bs, uint8_t* __restrict__ argmax, int accumulate,
                                  int32_t* __restrict__ assign, int B, int C, int Tf, int Hf,
                                  int Wf, int Tx, int K, int pst, float scale, int P) {
  const int box = blockIdx.x;  // (b, t, k) flattened
  const int k = box % K, t = (box / K) % Tx, b = box / (K * Tx);
  (void)k;
  const int slice = (Tf == 1) ? 0 : (Tx == 1 ? t : t / pst);
  if (threadIdx.x == 0 && assign) {
    assign[2 * box] = b;
    assign[2 * box + 1] = slice;
  }
  const float* bx = boxes + (int64_t)box * 4;
  const float x1 = bx[0] * scale - 0.5f, y1 = bx[1] * scale - 0.5f;
  const float x2 = bx[2] * scale - 0.5f, y2 = bx[3] * scale - 0.5f;
  const float rw = x2 - x1, rh = y2 - y1;
  const float bw = rw / (float)P, bh = rh / (float)P;
  const int gh = (int)ceilf(rh / (float)P), gw = (int)ceilf(rw / (float)P);
  const T* fm = feat + (int64_t)b * feat_bs + (1 + (int64_t)slice * Hf * Wf) * C;
  T* trow = tokens + (int64_t)b * tok_bs + (int64_t)(box - b * Tx * K) * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float best = -INFINITY;
    int arg = 0;
    for (int ph = 0; ph < P; ++ph)
      for (int pw = 0; pw < P; ++pw) {
        const float v = roi_bin(fm, C, c, Hf, Wf, y1, x1, bh, bw, gh, gw, ph, pw);
        if (v > best) {  // first maximal bin wins (torch.max semantics)
          best = v;
          arg = ph * P + pw;
        }
      }
    if (argmax) argmax[(int64_t)box * C + c] = (uint8_t)arg;
    trow[c] = from_f<T>(accumulate ? best + to_f(trow[c]) : best);
  }
}

// Backward of roi_tokens: the gradient of token (box, c) goes to the samples of its arg-max bin, spread over the four
// bilinear taps of every sample (weight / sample count), accumulated with fp32 atomics into dfeat [B, Tf*Hf*Wf, C].
template <typename T>
__global__ void roi_tokens_bwd_kernel(const T* __restrict__ dtok, int64_t dtok_bs, const uint8_t* __restrict__ argmax,
                                      const float* __restrict__ boxes, float* __restrict__ dfeat, int B, int C, int Tf,
                                      int Hf, int Wf, int Tx, int K, int pst, float scale, int P) {
  const int box = blockIdx.x;
  const int t = (box / K) % Tx, b = box / (K * Tx);
  const int slice = (Tf == 1) ? 0 : (Tx == 1 ? t : t / pst);
  const float* bx = boxes + (int64_t)box * 4;
  const float x1 = bx[0] * scale - 0.5f, y1 = bx[1] * scale - 0.5f;
  const float x2 = bx[2] * scale - 0.5f, y2 = bx[3] * scale - 0.5f;
  const float rw = x2 - x1, rh = y2 - y1;
  const float bw = rw / (float)P, bh = rh / (float)P;
  const int gh = (int)ceilf(rh / (float)P), gw = (int)ceilf(rw / (float)P);
  const int cnt = gh * gw;
  if (cnt <= 0) return;  // no samples: the token is the constant 0
  const float inv = 1.f / (float)cnt;
  float* df = dfeat + ((int64_t)b * Tf + slice) * Hf * Wf * C;
  const T* grow = dtok + (int64_t)b * dtok_bs + (int64_t)(box - b * Tx * K) * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float g = to_f(grow[c]) * inv;
    if (g == 0.f) continue;
    const int bin = argmax[(int64_t)box * C + c];
    const int ph = bin / P, pw = bin - ph * P;
    for (int iy = 0; iy < gh; ++iy) {
      const float y = y1 + ph * bh + (iy + 0.5f) * bh / (float)gh;
      for (int ix = 0; ix < gw; ++ix) {
        const float x = x1 + pw * bw + (ix + 0.5f) * bw / (float)gw;
        int yl, xl, yh, xh;
        float w1, w2, w3, w4;
        bool valid;
        bilinear_setup(y, x, Hf, Wf, yl, xl, yh, xh, w1, w2, w3, w4, valid);
        if (!valid) continue;
        atomicAdd(df + ((int64_t)yl * Wf + xl) * C + c, g * w1);
        atomicAdd(df + ((int64_t)yl * Wf + xh) * C + c, g * w2);
        atomicAdd(df + ((int64_t)yh * Wf + xl) * C + c, g * w3);
        atomicAdd(df + ((int64_t)yh * Wf + xh) * C + c, g * w4);
      }
    }
  }
}

// bf16, C % 8 == 0: same arithmetic, restructured for the memory system.
//   * the 1-D interpolation set-up (clamps, neighbour indices, weights, validity: separable in y and x) is computed once per
//     box into shared memory instead of once per channel thread and sample;
//   * a thread owns 8 consecutive channels (16-byte loads of the token-major feature rows) and the 256 threads of the CTA
//     split into 256 / (C/8) groups that take the 49 bins round-robin; the per-bin maxima are combined through smem.
struct Samp1D {
  int lo, hi;
  float wlo, whi;  // both 0 for an invalid (out-of-range) sample
};
constexpr int ROI_MAXS = 512;

__device__ __forceinline__ Samp1D samp1d(float v, int n) {
  Samp1D s;
  s.lo = s.hi = 0;
  s.wlo = s.whi = 0.f;
  if (v < -1.0f || v > (float)n) return s;
  if (v <= 0.f) v = 0.f;
  int lo = (int)v, hi;
  if (lo >= n - 1) { hi = lo = n - 1; v = (float)lo; } else { hi = lo + 1; }
  const float l = v - lo;
  s.lo = lo; s.hi = hi; s.wlo = 1.f - l; s.whi = l;
  return s;
}

__global__ void __launch_bounds__(256) roi_tokens_vec8_kernel(const bf16* __restrict__ feat, int64_t feat_bs,
                                                              const float* __restrict__ boxes, bf16* __restrict__ tokens,
                                                              int64_t tok_bs, uint8_t* __restrict__ argmax, int accumulate,
                                                              int32_t* __restrict__ assign, int B, int C, int Tf, int Hf,
                                                              int Wf, int Tx, int K, int pst, float scale, int P) {
  __shared__ Samp1D ys[ROI_MAXS], xs[ROI_MAXS];
  __shared__ float red[256 * 8];
  __shared__ uint8_t redarg[256 * 8];
  const int box = blockIdx.x;  // (b, t, k) flattened
  const int t = (box / K) % Tx, b = box / (K * Tx);
  const int slice = (Tf == 1) ? 0 : (Tx == 1 ? t : t / pst);
  if (threadIdx.x == 0 && assign) {
    assign[2 * box] = b;
    assign[2 * box + 1] = slice;
  }
  const float* bx = boxes + (int64_t)box * 4;
  const float x1 = bx[0] * scale - 0.5f, y1 = bx[1] * scale - 0.5f;
  const float x2 = bx[2] * scale - 0.5f, y2 = bx[3] * scale - 0.5f;
  const float rw = x2 - x1, rh = y2 - y1;
  const float bw = rw / (float)P, bh = rh / (float)P;
  const int gh = (int)ceilf(rh / (float)P), gw = (int)ceilf(rw / (float)P);
  // boxes far larger than the frame (more than ROI_MAXS 1-D samples) compute the set-up on the fly instead
  const bool tab = (int64_t)P * gh <= ROI_MAXS && (int64_t)P * gw <= ROI_MAXS;
  if (tab) {
    for (int i = threadIdx.x; i < P * gh; i += blockDim.x) {
      const int ph = i / gh, iy = i - ph * gh;
      ys[i] = samp1d(y1 + ph * bh + (iy + 0.5f) * bh / (float)gh, Hf);
    }
    for (int i = threadIdx.x; i < P * gw; i += blockDim.x) {
      const int pw = i / gw, ix = i - pw * gw;
      xs[i] = samp1d(x1 + pw * bw + (ix + 0.5f) * bw / (float)gw, Wf);
    }
  }
  __syncthreads();
  const int ct = C >> 3;                 // channel threads
  const int ngrp = blockDim.x / ct;      // bin groups
  const int cthr = threadIdx.x % ct, grp = threadIdx.x / ct;
  const bf16* fm = feat + (int64_t)b * feat_bs + (1 + (int64_t)slice * Hf * Wf) * C + cthr * 8;
  const int cnt = gh * gw;
  const float inv = 1.f / (float)(cnt > 0 ? cnt : 1);
  float best[8];
  int barg[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    best[u] = -INFINITY;
    barg[u] = 0;
  }
  auto fma8 = [](float acc[8], const uint4& v, float w) {
    const uint32_t q[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      acc[2 * u] = fmaf(w, __uint_as_float(q[u] << 16), acc[2 * u]);
      acc[2 * u + 1] = fmaf(w, __uint_as_float(q[u] & 0xffff0000u), acc[2 * u + 1]);
    }
  };
  if (grp < ngrp) {
    for (int bin = grp; bin < P * P; bin += ngrp) {
      const int ph = bin / P, pw = bin - ph * P;
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int iy = 0; iy < gh; ++iy) {
        const Samp1D sy = tab ? ys[ph * gh + iy] : samp1d(y1 + ph * bh + (iy + 0.5f) * bh / (float)gh, Hf);
        if (sy.wlo == 0.f && sy.whi == 0.f) continue;
        const bf16* r0 = fm + (int64_t)sy.lo * Wf * C;
        const bf16* r1 = fm + (int64_t)sy.hi * Wf * C;
        for (int ix = 0; ix < gw; ++ix) {
          const Samp1D sx = tab ? xs[pw * gw + ix] : samp1d(x1 + pw * bw + (ix + 0.5f) * bw / (float)gw, Wf);
          if (sx.wlo == 0.f && sx.whi == 0.f) continue;
          const uint4 v00 = __ldg(reinterpret_cast<const uint4*>(r0 + sx.lo * C));
          const uint4 v01 = __ldg(reinterpret_cast<const uint4*>(r0 + sx.hi * C));
          const uint4 v10 = __ldg(reinterpret_cast<const uint4*>(r1 + sx.lo * C));
          const uint4 v11 = __ldg(reinterpret_cast<const uint4*>(r1 + sx.hi * C));
          fma8(acc, v00, sy.wlo * sx.wlo);
          fma8(acc, v01, sy.wlo * sx.whi);
          fma8(acc, v10, sy.whi * sx.wlo);
          fma8(acc, v11, sy.whi * sx.whi);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float v = acc[u] * inv;
        if (v > best[u]) {  // bins arrive in increasing order inside a group: the first maximal bin is kept
          best[u] = v;
          barg[u] = bin;
        }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    red[threadIdx.x * 8 + u] = best[u];
    redarg[threadIdx.x * 8 + u] = (uint8_t)barg[u];
  }
  __syncthreads();
  if (threadIdx.x < ct) {
    const int ng = ngrp < P * P ? ngrp : P * P;  // groups that owned at least one bin
    float m[8];
    int ma[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      m[u] = red[threadIdx.x * 8 + u];
      ma[u] = redarg[threadIdx.x * 8 + u];
    }
    for (int g2 = 1; g2 < ng; ++g2)
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float v = red[(g2 * ct + threadIdx.x) * 8 + u];
        const int a = redarg[(g2 * ct + threadIdx.x) * 8 + u];
        if (v > m[u] || (v == m[u] && a < ma[u])) {  // ties: the lowest bin index (first maximal bin)
          m[u] = v;
          ma[u] = a;
        }
      }
    bf16* trow = tokens + (int64_t)b * tok_bs + (int64_t)(box - b * Tx * K) * C + threadIdx.x * 8;
    if (accumulate) {
      const uint4 old = *reinterpret_cast<const uint4*>(trow);
      const uint32_t q[4] = {old.x, old.y, old.z, old.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        m[2 * u] += __uint_as_float(q[u] << 16);
        m[2 * u + 1] += __uint_as_float(q[u] & 0xffff0000u);
      }
    }
    __nv_bfloat162 h2[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) h2[u] = __floats2bfloat162_rn(m[2 * u], m[2 * u + 1]);
    *reinterpret_cast<uint4*>(trow) = *reinterpret_cast<uint4*>(h2);
    if (argmax) {
      uint8_t a8[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) a8[u] = (uint8_t)ma[u];
      *reinterpret_cast<uint2*>(argmax + (int64_t)box * C + threadIdx.x * 8) = *reinterpret_cast<uint2*>(a8);
    }
  }
}

template <typename T>
__global__ void roi_align_kernel(const T* __restrict__ feat, const float* __restrict__ rois, T* __restrict__ out, int N,
                                 int C, int H, int W, int R, int P, float scale, int sampling, int aligned) {
  const int r = blockIdx.x / (P * P);
  const int bin = blockIdx.x % (P * P);
  const int ph = bin / P, pw = bin % P;
  const float* roi = rois + (int64_t)r * 5;
  const int n = (int)roi[0];
  const float off = aligned ? 0.5f : 0.f;
  const float x1 = roi[1] * scale - off, y1 = roi[2] * scale - off;
  const float x2 = roi[3] * scale - off, y2 = roi[4] * scale - off;
  float rw = x2 - x1, rh = y2 - y1;
  if (!aligned) { rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); }
  const float bw = rw / (float)P, bh = rh / (float)P;
  const int gh = sampling > 0 ? sampling : (int)ceilf(rh / (float)P);
  const int gw = sampling > 0 ? sampling : (int)ceilf(rw / (float)P);
  const T* fm = feat + (int64_t)n * H * W * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x)
    out[((int64_t)blockIdx.x) * C + c] = from_f<T>(roi_bin(fm, C, c, H, W, y1, x1, bh, bw, gh, gw, ph, pw));
}

// utils/box_ops.py:140-194.  One thread per sample; boxes [n,4,4] (hand0, hand1, obj0, obj1).
__global__ void match_haog_kernel(float* __restrict__ boxes, int64_t* __restrict__ contact, int64_t n) {
  const float HIGH = 1e8f;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (int64_t)gridDim.x * blockDim.x) {
    float* bx = boxes + s * 16;
    float v[4][4];
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) v[i][j] = bx[i * 4 + j];
    float cost[2][2];
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 2; ++j) {
        float dx = __fsub_rn(v[i][0], v[2 + j][0]), dy = __fsub_rn(v[i][1], v[2 + j][1]);
        cost[i][j] = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
      }
    bool zero[4];
    for (int i = 0; i < 4; ++i) zero[i] = v[i][0] == 0.f && v[i][1] == 0.f && v[i][2] == 0.f && v[i][3] == 0.f;
    // cost[:, obj_is_zero] = HIGH ; cost[:, hand_is_zero] = HIGH  (both index COLUMNS, as the reference does)
    for (int j = 0; j < 2; ++j)
      if (zero[2 + j] || zero[j]) cost[0][j] = cost[1][j] = HIGH;
    float ord1 = __fadd_rn(cost[0][0], cost[1][1]), ord2 = __fadd_rn(cost[0][1], cost[1][0]);
    float c1, c2;
    if (ord2 < ord1) {
      const int perm[4] = {0, 2, 3, 1};
      for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) bx[i * 4 + j] = v[perm[i]][j];
      c1 = cost[0][1];
      c2 = cost[1][0];
    } else {
      c1 = cost[0][0];
      c2 = cost[1][1];
    }
    contact[2 * s] = c1 == HIGH ? -1 : (c1 < 0.1f ? 3 : 0);
    contact[2 * s + 1] = c2 == HIGH ? -1 : (c2 < 0.1f ? 3 : 0);
  }
}

__global__ void zero_empty_boxes_kernel(float* __restrict__ b, int64_t n, float eps) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float w = b[4 * i + 2], h = b[4 * i + 3];
    if (w <= eps || h <= eps) b[4 * i] = b[4 * i + 1] = b[4 * i + 2] = b[4 * i + 3] = 0.f;
  }
}

extern "C" {

int svit_roi_tokens_fwd(const void* feat, int64_t feat_batch_stride, const float* boxes, void* tokens,
                        int64_t tokens_batch_stride, uint8_t* argmax, int accumulate, int32_t* assign, int B, int C, int Tf,
                        int Hf, int Wf, int Tx, int K, int patch_stride_t, float spatial_scale, int P, int dtype,
                        void* stream) {
  if (B < 0 || K < 0 || Tx < 1 || P < 1 || P > 15 || patch_stride_t < 1 || Tf < 1) return SVIT_EINVAL;
  if (Tf > 1 && Tx > 1 && (Tx - 1) / patch_stride_t >= Tf) return SVIT_EINVAL;
  int64_t nbox = (int64_t)B * Tx * K;
  if (nbox == 0) return 0;
  const int64_t tbs = tokens_batch_stride > 0 ? tokens_batch_stride : (int64_t)Tx * K * C;
  cudaStream_t st = (cudaStream_t)stream;
  int threads = C >= 256 ? 256 : (C >= 128 ? 128 : 96);
  if (dtype == SVIT_F32)
    roi_tokens_kernel<float><<<(unsigned)nbox, threads, 0, st>>>((const float*)feat, feat_batch_stride, boxes, (float*)tokens, tbs, argmax, accumulate, assign, B, C, Tf, Hf, Wf, Tx, K, patch_stride_t, spatial_scale, P);
  else if (dtype == SVIT_BF16 && C % 8 == 0 && C <= 2048 && feat_batch_stride % 8 == 0 && tbs % 8 == 0 &&
           ((reinterpret_cast<uintptr_t>(feat) | reinterpret_cast<uintptr_t>(tokens)) & 15) == 0 &&
           (!argmax || (reinterpret_cast<uintptr_t>(argmax) & 7) == 0))
    roi_tokens_vec8_kernel<<<(unsigned)nbox, 256, 0, st>>>((const bf16*)feat, feat_batch_stride, boxes, (bf16*)tokens, tbs, argmax, accumulate, assign, B, C, Tf, Hf, Wf, Tx, K, patch_stride_t, spatial_scale, P);
  else if (dtype == SVIT_BF16)
    roi_tokens_kernel<bf16><<<(unsigned)nbox, threads, 0, st>>>((const bf16*)feat, feat_batch_stride, boxes, (bf16*)tokens, tbs, argmax, accumulate, assign, B, C, Tf, Hf, Wf, Tx, K, patch_stride_t, spatial_scale, P);
  else
    return SVIT_EINVAL;
  SVIT_CHECK_LAUNCH();
  return 0;
}

int svit_roi_tokens_bwd(const void* dtokens, int64_t dtokens_batch_stride, const uint8_t* argmax, const float* boxes,
                        float* dfeat, int B, int C, int Tf, int Hf, int Wf, int Tx, int K, int patch_stride_t,
                        float spatial_scale, int P, int dtype, void* stream) {
  if (B < 0 || K < 0 || Tx < 1 || P < 1 || P > 15 || patch_stride_t < 1 || Tf < 1 || !argmax || !dfeat) return SVIT_EINVAL;
  if (Tf > 1 && Tx > 1 && (Tx - 1) / patch_stride_t >= Tf) return SVIT_EINVAL;
  int64_t nbox = (int64_t)B * Tx * K;
  if (nbox == 0) return 0;
  const int64_t tbs = dtokens_batch_stride > 0 ? dtokens_batch_stride : (int64_t)Tx * K * C;
  cudaStream_t st = (cudaStream_t)stream;
  int threads = C >= 256 ? 256 : (C >= 128 ? 128 : 96);
  if (dtype == SVIT_F32)
    roi_tokens_bwd_kernel<float><<<(unsigned)nbox, threads, 0, st>>>((const float*)dtokens, tbs, argmax, boxes, dfeat, B, C, Tf, Hf, Wf, Tx, K, patch_stride_t, spatial_scale, P);
  else if (dtype == SVIT_BF16)
    roi_tokens_bwd_kernel<bf16><<<(unsigned)nbox, threads, 0, st>>>((const bf16*)dtokens, tbs, argmax, boxes, dfeat, B, C, Tf, Hf, Wf, Tx, K, patch_stride_t, spatial_scale, P);
  else
    return SVIT_EINVAL;
  SVIT_CHECK_LAUNCH();
  return 0;
}

int svit_roi_align_fwd(const void* feat, const float* rois, void* out, int N, int C, int H, int W, int R, int P,
                       float spatial_scale, int sampling_ratio, int aligned, int dtype, void* stream) {
  if (R < 0 || P < 1 || N < 1) return SVIT_EINVAL;
  if (R == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int threads = C >= 256 ? 256 : 128;
  unsigned grid = (unsigned)((int64_t)R * P * P);
  if (dtype == SVIT_F32)
    roi_align_kernel<float><<<grid, threads, 0, st>>>((const float*)feat, rois, (float*)out, N, C, H, W, R, P, spatial_scale, sampling_ratio, aligned);
  else if (dtype == SVIT_BF16)
    roi_align_kernel<bf16><<<grid, threads, 0, st>>>((const bf16*)feat, rois, (bf16*)out, N, C, H, W, R, P, spatial_scale, sampling_ratio, aligned);
  else
    return SVIT_EINVAL;
  SVIT_CHECK_LAUNCH();
  return 0;
}

int svit_match_haog(float* boxes, int64_t* contact, int64_t n, void* stream) {
  if (n < 0) return SVIT_EINVAL;
  if (n == 0) return 0;
  match_haog_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(boxes, contact, n);
  SVIT_CHECK_LAUNCH();
  return 0;
}

int svit_zero_empty_boxes(float* boxes_cxcywh, int64_t n, float eps, void* stream) {
  if (n < 0) return SVIT_EINVAL;
  if (n == 0) return 0;
  zero_empty_boxes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(boxes_cxcywh, n, eps);
  SVIT_CHECK_LAUNCH();
  return 0;
}

int svit_abi_version(void) { return 100; }

}  // extern "C"
