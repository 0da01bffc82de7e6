// attention_pool forward (conv + object-token scale + LayerNorm), bf16 production kernels.
// Reference: slowfast/models/attention.py:13-65.  HBM-bound: every input token slice is read once from
// DRAM (halo re-reads are L2 hits) and every output token written once.
//
//   pool_ln_march_kernel<S, SPR>  stride (1,S,S), S = 1 or 2: persistent CTAs; thread = (strip of outputs along W, one
//                                 bf16x2 channel pair); the CTA marches over T through a TMA ring of input planes.
//   pool_ln_direct_kernel         stride >= 4 (windows do not overlap) or unaligned input: a half-warp per output token
//                                 (lane l16 owns channel words l16, l16+16, l16+32), taps straight from global / L2.
#include <cstdlib>
#include <type_traits>

#include "tc_common.cuh"

namespace {

constexpr int PD = 96;
constexpr int TAPS = 27;

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

// two fp32 FMAs per instruction (sm_100 FFMA2): acc = x * w + acc on both halves of a 64-bit register pair
__device__ __forceinline__ void fma2(float2& acc, const float2 x, const float2 w) {
#ifdef POOL_SCALAR_FMA
  acc.x = fmaf(x.x, w.x, acc.x);
  acc.y = fmaf(x.y, w.y, acc.y);
  return;
#endif
  unsigned long long a = *reinterpret_cast<unsigned long long*>(&acc);
  const unsigned long long xx = *reinterpret_cast<const unsigned long long*>(&x);
  const unsigned long long ww = *reinterpret_cast<const unsigned long long*>(&w);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a) : "l"(xx), "l"(ww));
  acc = *reinterpret_cast<float2*>(&a);
}

__device__ __forceinline__ float2 shfl_xor2(float2 v, int m) {
  return make_float2(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
}

// packed fp32x2 helpers (FADD2 / FMUL2 / FFMA2)
__device__ __forceinline__ float2 padd2(const float2 a, const float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<const unsigned long long*>(&a)),
      "l"(*reinterpret_cast<const unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 pmul2(const float2 a, const float2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<const unsigned long long*>(&a)),
      "l"(*reinterpret_cast<const unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 pfma2(const float2 a, const float2 b, const float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<const unsigned long long*>(&a)),
      "l"(*reinterpret_cast<const unsigned long long*>(&b)), "l"(*reinterpret_cast<const unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}

constexpr int MK_PITCH = 50;  // staging token pitch in float2: conflict-free 16-byte reads by the LN lanes

// LayerNorm + store of the first `ntok` staged tokens (pre-LN fp32, MK_PITCH float2 per token); four lanes per token,
// 24 channels per lane, packed fp32x2 arithmetic.
__device__ __forceinline__ void mk_ln_flush(const float2* stg, const long long* stg_tok, const float* aff, int ntok, float eps,
                                            bf16* __restrict__ obase, int tid, bf16* __restrict__ pre_base = nullptr) {
  const int tok = tid >> 2, q = tid & 3;
  if ((tid & ~31) >= ntok * 4) return;  // whole warp idle
  const bool live = tok < ntok;
  const float2* src = stg + (live ? tok : 0) * MK_PITCH + q * 12;
  float2 x[12];
#pragma unroll
  for (int i = 0; i < 12; i += 2) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    x[i] = make_float2(v.x, v.y);
    x[i + 1] = make_float2(v.z, v.w);
  }
  float2 s2 = padd2(x[0], x[1]);
#pragma unroll
  for (int i = 2; i < 12; ++i) s2 = padd2(s2, x[i]);
  float sum = s2.x + s2.y;
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  sum += __shfl_xor_sync(0xffffffffu, sum, 2);
  const float nmean = sum * (-1.f / PD);
  const float2 nm2 = make_float2(nmean, nmean);
  const long long gtok = live ? stg_tok[tok] : -1;
  if (pre_base && gtok >= 0) {  // training: the pre-LayerNorm row for the backward
    uint32_t pw[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) pw[i] = pack2(x[i].x, x[i].y);
    uint4* pd = reinterpret_cast<uint4*>(pre_base + gtok * PD + q * 24);
    pd[0] = make_uint4(pw[0], pw[1], pw[2], pw[3]);
    pd[1] = make_uint4(pw[4], pw[5], pw[6], pw[7]);
    pd[2] = make_uint4(pw[8], pw[9], pw[10], pw[11]);
  }
  float2 q2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    x[i] = padd2(x[i], nm2);
    q2 = pfma2(x[i], x[i], q2);
  }
  float var = q2.x + q2.y;
  var += __shfl_xor_sync(0xffffffffu, var, 1);
  var += __shfl_xor_sync(0xffffffffu, var, 2);
  const float rstd = rsqrtf(var * (1.f / PD) + eps);
  if (gtok < 0) return;
  const float2 r2 = make_float2(rstd, rstd);
  const float4* gm = reinterpret_cast<const float4*>(aff + q * 24);
  const float4* bt = reinterpret_cast<const float4*>(aff + PD + q * 24);
  uint32_t o[12];
#pragma unroll
  for (int i = 0; i < 12; i += 2) {
    const float4 g4 = gm[i >> 1], b4 = bt[i >> 1];
    const float2 y0 = pfma2(pmul2(x[i], r2), make_float2(g4.x, g4.y), make_float2(b4.x, b4.y));
    const float2 y1 = pfma2(pmul2(x[i + 1], r2), make_float2(g4.z, g4.w), make_float2(b4.z, b4.w));
    o[i] = pack2(y0.x, y0.y);
    o[i + 1] = pack2(y1.x, y1.y);
  }
  uint4* dst = reinterpret_cast<uint4*>(obase + gtok * PD + q * 24);
  dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
  dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
  dst[2] = make_uint4(o[8], o[9], o[10], o[11]);
}


#ifdef SVIT_TIMELINE
__device__ unsigned long long* g_pool_dbg = nullptr;
#define PTL(tag)                                                              \
  do {                                                                        \
    if (g_pool_dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0 && tl_n < 1024) { \
      g_pool_dbg[tl_n * 2] = (unsigned long long)(tag);                       \
      g_pool_dbg[tl_n * 2 + 1] = (unsigned long long)clock64();               \
      ++tl_n;                                                                 \
    }                                                                         \
  } while (0)
#else
#define PTL(tag) do { } while (0)
#endif

// ------------------------------------------------------------------------------------------------ march
// pool_ln_march_kernel<S, SPR>: thread = (strip of SW consecutive outputs along W, one bf16x2 channel pair); 384 threads =
// 8 strips x 48 pairs; 2 CTAs per SM.  A work item ("column") is one output tile of one (batch, head) over all T planes:
// the CTA marches over the input planes and every plane is read from shared memory exactly once per strip: its three
// input rows feed the accumulators of the three output planes t-1, t, t+1 (three rotating register sets), so one LDS.32 +
// 2 conversions feed up to 27 FFMA2.  A finished output plane is normalised straight from the accumulators (ln_plane).
// CTAs are PERSISTENT: grid = min(columns, 2 x SMs), the tap weights are staged once per CTA, and the TMA plane ring
// (one box per plane, two planes ahead) runs across column boundaries, so the next column's first planes are in flight
// while the current column finishes (a per-column prologue costs ~20 % of a column at T = 8: measured by merging two
// columns into one CTA).  The cls / object rows of every (batch, head) are separate, cheap items done after the columns.
// SPR = strips per output row: 2 (tile 4 rows x 2 strips) or 1 (8 rows x 1 strip: 7-wide grids waste no strips).
constexpr int MK_THREADS = 384, MK_STRIPS = 8, MK_SLOTS = 3;
template <int S, int SPR> struct MkCfg {
  static_assert(S == 1 || SPR == 2, "stride 2 uses two strips per row");
  static constexpr int SW = S == 1 ? 7 : 4;
  static constexpr int ROWS = MK_STRIPS / SPR;
  static constexpr int TW = S == 1 ? SPR * SW : 7;      // S = 2: the second strip's 4th output is masked
  static constexpr int IW = S == 1 ? TW + 2 : 15;       // S = 2: strip 2 over-reads 2 tokens (they only reach the masked output)
  static constexpr int IH = (ROWS - 1) * S + 3;
  static constexpr int XN = (SW - 1) * S + 3;
  static constexpr int PLANE_BYTES = IH * IW * PD * 2;  // one TMA box
  static constexpr int SLOT_WORDS = ((PLANE_BYTES + 127) / 128) * 32 + (S == 2 ? 128 : 0);  // 128-byte aligned (+ over-read room)
  static constexpr int RING_BYTES = MK_SLOTS * SLOT_WORDS * 4;
  static constexpr int NTOK = MK_STRIPS * SW;           // cls / object rows staged per flush (aliases the idle ring)
  static_assert(NTOK * MK_PITCH * 8 + NTOK * 8 <= RING_BYTES, "staging tile must fit into the plane ring");
  static constexpr int OFF_AFF = RING_BYTES;                    // gamma | beta | w_eff
  static constexpr int OFF_W = OFF_AFF + 3 * PD * 4;            // tap weights as float2 [27][48]
  static constexpr int OFF_BAR = OFF_W + TAPS * (PD / 2) * 8;   // one mbarrier per ring slot
  static constexpr int OFF_RED = OFF_BAR + 64;                  // LayerNorm partial sums [2][24 groups][16]
  static constexpr int SMEM = OFF_RED + 2 * 24 * 16 * 4;
};

// SAVE (training): the pre-LayerNorm rows are written to `pre` as well (same layout as out), so that the backward does
// not have to recompute the convolution.
template <int S, int SPR, bool PERSIST, bool SAVE>
__global__ void __launch_bounds__(MK_THREADS, 2)
pool_ln_march_kernel(const __grid_constant__ CUtensorMap tmap, const bf16* __restrict__ in, Geom g, const float* __restrict__ w,
                     const float* __restrict__ frac, const float* __restrict__ gamma, const float* __restrict__ beta,
                     bf16* __restrict__ out, float eps, int tiles, int ncols, bf16* __restrict__ pre) {
  using C = MkCfg<S, SPR>;
  constexpr int IW = C::IW, XN = C::XN, SW = C::SW;
  constexpr int NV = SW > 4 ? 16 : 8;  // LayerNorm butterfly width: token sums in v[0, NV/2), sums of squares in v[NV/2, NV)
  extern __shared__ __align__(128) unsigned char smem[];
  uint32_t* ring = reinterpret_cast<uint32_t*>(smem);
  float* aff = reinterpret_cast<float*>(smem + C::OFF_AFF);
  float2* sw2 = reinterpret_cast<float2*>(smem + C::OFF_W);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  float* red = reinterpret_cast<float*>(smem + C::OFF_RED);
  const int tiles_w = (g.Wo + C::TW - 1) / C::TW;
  const int64_t Lo = (int64_t)g.T * g.Ho * g.Wo, L = (int64_t)g.T * g.H * g.W;
  const int64_t Nout = 1 + Lo + g.O;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmap);
    for (int i = 0; i < MK_SLOTS; ++i) tc::mbar_init(&full[i], 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  // plane m of this CTA's sequence = plane (m % T) of its (m / T)-th column; thread 0 walks (column, plane, slot) two planes
  // ahead; the walker lives in shared memory (registers are the scarce resource of this kernel)
  int* pf = reinterpret_cast<int*>(smem + C::OFF_BAR + 32);  // {column, plane, slot, c0 = head*96, c1 = iw0, c2 = ih0, c4 = b}
  auto pf_column = [&](int col) {  // TMA coordinates of a column: once per column, not per plane
    const int tile = col % tiles, bh = col / tiles;
    pf[0] = col;
    pf[3] = (bh % g.h) * PD;
    pf[4] = (tile % tiles_w) * C::TW * S - 1;
    pf[5] = (tile / tiles_w) * C::ROWS * S - 1;
    pf[6] = bh / g.h;
  };
  // !PERSIST (one column per CTA): the coordinates stay in registers and the ring slot is the compile-time R of step()
  const int c_tile = blockIdx.x % tiles, c_bh = blockIdx.x / tiles;
  const int c_w = (c_tile % tiles_w) * C::TW * S - 1, c_h = (c_tile / tiles_w) * C::ROWS * S - 1;
  auto load_plane = [&](int t) {  // thread 0 only, !PERSIST
    const int sl = t % MK_SLOTS;
    tc::mbar_arrive_expect_tx(&full[sl], C::PLANE_BYTES);
    tc::tma_load_5d(ring + sl * C::SLOT_WORDS, &tmap, &full[sl], (c_bh % g.h) * PD, c_w, c_h, t, c_bh / g.h);
  };
  auto load_next = [&]() {  // thread 0 only, PERSIST
    const int pf_col = pf[0], pf_t = pf[1], pf_slot = pf[2];
    if (pf_col < ncols) {
      tc::mbar_arrive_expect_tx(&full[pf_slot], C::PLANE_BYTES);
      tc::tma_load_5d(ring + pf_slot * C::SLOT_WORDS, &tmap, &full[pf_slot], pf[3], pf[4], pf[5], pf_t, pf[6]);
      pf[2] = pf_slot == MK_SLOTS - 1 ? 0 : pf_slot + 1;
      if (pf_t + 1 == g.T) {
        pf[1] = 0;
        if (pf_col + (int)gridDim.x < ncols) pf_column(pf_col + gridDim.x);
        else pf[0] = ncols;
      } else {
        pf[1] = pf_t + 1;
      }
    }
  };
  if (threadIdx.x == 0) {
    if (PERSIST) {
      pf_column(blockIdx.x);
      pf[1] = 0;
      pf[2] = 0;
      load_next();
      load_next();
    } else {
      for (int t = 0; t < 2 && t < g.T; ++t) load_plane(t);
    }
  }
  for (int i = threadIdx.x; i < PD; i += MK_THREADS) {
    aff[i] = gamma[i];
    aff[PD + i] = beta[i];
  }
  for (int i = threadIdx.x; i < TAPS * 48; i += MK_THREADS) {
    const int tp = i / 48, pr = i - tp * 48;
    sw2[i] = make_float2(__ldg(w + (2 * pr) * TAPS + tp), __ldg(w + (2 * pr + 1) * TAPS + tp));
  }

  const int strip = threadIdx.x / 48, wd = threadIdx.x - strip * 48;   // strip, channel-pair word
  const int srow = strip / SPR, scol = (strip % SPR) * SW;
  const int lane = threadIdx.x & 31;
  const int plane_words = g.Ho * g.Wo * (PD / 2);
  const uint32_t* xbase = ring + ((srow * S) * IW + scol * S) * 48 + wd;
  const float2* wbase = sw2 + wd;
  const float2 gam = make_float2(__ldg(gamma + 2 * wd), __ldg(gamma + 2 * wd + 1));
  const float2 bet = make_float2(__ldg(beta + 2 * wd), __ldg(beta + 2 * wd + 1));
  // LayerNorm statistics meet in red[plane parity][16-lane group (24)][16]; a strip = groups 3*strip .. 3*strip + 2
  const int l16 = lane & 15;
  const int my_tok = l16 & (NV / 2 - 1);                        // the token whose statistics this lane finishes
  const int red_wr = (threadIdx.x >> 4) * 16 + l16;
  const int red_rd = strip * 48 + (NV == 16 ? my_tok : 2 * my_tok);
  const int bar_id = 1 + (strip >> 1);
  // per column: this thread's output tokens are row ho0 + srow, columns wo0 + scol + o for o < nvalid; optr walks the planes
  int nvalid = 0;
  uint32_t* optr = nullptr;
  uint32_t* pptr = nullptr;  // SAVE: walks `pre` like optr walks `out`
  int slot = 0, parity = 0;  // ring position of the plane being consumed

  float2 acc[3][SW];

  // LayerNorm of a finished output plane straight from the accumulators.  The 96 channels of a token are the 48
  // threads of its strip = three 16-lane groups.  Each thread contributes (x + y, x^2 + y^2) for its SW tokens; a
  // transposing butterfly over the 16 lanes of a group (NV - 1 shuffles) leaves lane l with the group total of
  // value l (NV = 16) or l / 2 (NV = 8); the three groups meet in shared memory behind a 96-thread named barrier
  // (two strips = three warps); lane o of every group then finishes token o (mean, rstd) and the 16 lanes fetch the
  // per-token scale / shift with two shuffles per token.  Stores: 4 bytes per thread and token, a warp writes
  // 128 contiguous bytes; the token rows of a strip are consecutive in memory (immediate offsets from optr).
  auto ln_plane = [&](float2 (&set)[SW], int buf) {
    float v[NV];
#pragma unroll
    for (int o = 0; o < NV / 2; ++o) {
      if (o < SW) {
        v[o] = set[o].x + set[o].y;
        v[NV / 2 + o] = fmaf(set[o].x, set[o].x, set[o].y * set[o].y);
      } else {
        v[o] = v[NV / 2 + o] = 0.f;
      }
    }
    {
      int d = 8;
#pragma unroll
      for (int hw = NV / 2; hw >= 1; hw >>= 1, d >>= 1) {
        const bool up = (lane & d) != 0;
#pragma unroll
        for (int i = 0; i < hw; ++i) {
          const float send = up ? v[i] : v[i + hw];
          const float keep = up ? v[i + hw] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, d);
        }
      }
      if (NV == 8) v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
    }
    float* rb = red + buf * (24 * 16);
    rb[red_wr] = v[0];
    asm volatile("bar.sync %0, 96;" ::"r"(bar_id) : "memory");
    const float* rr = rb + red_rd;
    const float sum = rr[0] + rr[16] + rr[32];
    const float sq = rr[8] + rr[24] + rr[40];
    const float mean = sum * (1.f / PD);
    const float var = fmaxf(fmaf(sq, 1.f / PD, -mean * mean), 0.f);
    const float sc = rsqrtf(var + eps);
    const float sh = -mean * sc;
    if (SAVE) {
#pragma unroll
      for (int o = 0; o < SW; ++o)
        if (o < nvalid) pptr[o * (PD / 2)] = pack2(set[o].x, set[o].y);
      pptr += plane_words;
    }
    uint32_t pk[SW];
#pragma unroll
    for (int o = 0; o < SW; ++o) {
      const float a = __shfl_sync(0xffffffffu, sc, (lane & 16) + o);
      const float c = __shfl_sync(0xffffffffu, sh, (lane & 16) + o);
      const float2 ag = make_float2(a * gam.x, a * gam.y);
      const float2 cg = make_float2(fmaf(c, gam.x, bet.x), fmaf(c, gam.y, bet.y));
      const float2 y = pfma2(set[o], ag, cg);
      pk[o] = pack2(y.x, y.y);
    }
    if (nvalid == SW) {
#pragma unroll
      for (int o = 0; o < SW; ++o) optr[o * (PD / 2)] = pk[o];
    } else {
#pragma unroll
      for (int o = 0; o < SW; ++o)
        if (o < nvalid) optr[o * (PD / 2)] = pk[o];
    }
    optr += plane_words;
  };

  int tl_n = 0;
  (void)tl_n;
  // One input plane tp of the current column: its three rows (kh) feed output planes tp+1 (kt = 0, a fresh accumulator
  // set: the first tap is a multiply, so sets are never zeroed), tp (kt = 1) and tp-1 (kt = 2, complete afterwards).
  // R = tp % 3 selects the rotating sets at compile time; E bit 0 = first plane of the column (no plane tp-1), bit 1 =
  // last plane (no plane tp+1; its own output plane is normalised here too).  No branches inside the tap loops.
  auto step = [&](auto rtag, auto etag, int tp) {
    constexpr int R = decltype(rtag)::value;
    constexpr int E = decltype(etag)::value;
    constexpr bool FIRST = (E & 1) != 0, LAST = (E & 2) != 0;
    PTL(100 + tp);
    __syncthreads();                           // every thread is done with the slot that the next load overwrites
    PTL(200 + tp);
    const uint32_t* pl;
    if (PERSIST) {
      if (threadIdx.x == 0) load_next();         // two planes ahead (possibly the next column's)
      tc::mbar_wait_hot(&full[slot], parity);    // this plane has landed
      pl = xbase + slot * C::SLOT_WORDS;
      if (++slot == MK_SLOTS) {
        slot = 0;
        parity ^= 1;
      }
    } else {
      if (threadIdx.x == 0 && tp + 2 < g.T) load_plane(tp + 2);
      tc::mbar_wait_hot(&full[R], (tp / MK_SLOTS) & 1);
      pl = xbase + R * C::SLOT_WORDS;
    }
    PTL(300 + tp);
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      float2 x[XN];
#pragma unroll
      for (int p = 0; p < XN; ++p) {
        const uint32_t v = pl[(kh * IW + p) * 48];
        x[p] = make_float2(lo_f(v), hi_f(v));
      }
#pragma unroll
      for (int kt = 0; kt < 3; ++kt) {
        if ((kt == 0 && LAST) || (kt == 2 && FIRST)) continue;  // temporal zero padding (compile time)
        float2 (&set)[SW] = acc[(R + 4 - kt) % 3];
        const bool fresh = kh == 0 && (kt == 0 || (kt == 1 && FIRST));
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float2 wt = wbase[((kt * 3 + kh) * 3 + kw) * 48];
#pragma unroll
          for (int o = 0; o < SW; ++o) {
            if (fresh && kw == 0) set[o] = pmul2(x[o * S], wt);
            else fma2(set[o], x[o * S + kw], wt);
          }
        }
      }
    }
    PTL(400 + tp);
    if (!FIRST) ln_plane(acc[(R + 2) % 3], (tp - 1) & 1);  // output plane tp-1 has now seen planes tp-2, tp-1, tp
    if (LAST) ln_plane(acc[R], tp & 1);                    // output plane T-1 (its t+1 neighbour is zero padding)
    PTL(600 + tp);
  };
  using R0 = std::integral_constant<int, 0>;
  using R1 = std::integral_constant<int, 1>;
  using R2 = std::integral_constant<int, 2>;
  using EMID = std::integral_constant<int, 0>;
  using EFIRST = std::integral_constant<int, 1>;
  using ELAST = std::integral_constant<int, 2>;
  using EONLY = std::integral_constant<int, 3>;

  for (int col = blockIdx.x; col < ncols; col += gridDim.x) {  // !PERSIST: exactly one iteration (grid = ncols)
    {
      const int tile = col % tiles, bh = col / tiles;
      const int wo0 = (tile % tiles_w) * C::TW, ho0 = (tile / tiles_w) * C::ROWS;
      nvalid = min(SW, min(C::TW - scol, g.Wo - (wo0 + scol)));
      if (ho0 + srow >= g.Ho || nvalid < 0) nvalid = 0;
      const int64_t row0 = ((int64_t)bh * Nout + 1 + (int64_t)(ho0 + srow) * g.Wo + wo0 + scol) * PD;
      optr = reinterpret_cast<uint32_t*>(out + row0) + wd;
      if (SAVE) pptr = reinterpret_cast<uint32_t*>(pre + row0) + wd;
    }
    if (g.T == 1) {
      step(R0{}, EONLY{}, 0);
    } else {
      step(R0{}, EFIRST{}, 0);
      int tp = 1;
      for (; tp + 3 <= g.T - 1; tp += 3) {
        step(R1{}, EMID{}, tp);
        step(R2{}, EMID{}, tp + 1);
        step(R0{}, EMID{}, tp + 2);
      }
      const int rem = g.T - 1 - tp;  // middle planes left before the last one: 0, 1 or 2
      if (rem == 0) {
        step(R1{}, ELAST{}, tp);
      } else if (rem == 1) {
        step(R1{}, EMID{}, tp);
        step(R2{}, ELAST{}, tp + 1);
      } else {
        step(R1{}, EMID{}, tp);
        step(R2{}, EMID{}, tp + 1);
        step(R0{}, ELAST{}, tp + 2);
      }
    }
  }

  // cls + object rows: item = one (batch, head), through an fp32 staging tile (aliases the idle ring) and the
  // 4-lanes-per-token LayerNorm.  Items are dealt from the last CTA backwards: those CTAs have the fewest columns.
  {
    float2* stg = reinterpret_cast<float2*>(ring);
    long long* stg_tok = reinterpret_cast<long long*>(smem + C::NTOK * MK_PITCH * 8);
    float* sweff = aff + 2 * PD;
    const int nbh = g.B * g.h;
    bool have_weff = false;
    for (int item = gridDim.x - 1 - blockIdx.x; item < nbh; item += gridDim.x) {
      if (!have_weff) {
        if (threadIdx.x < 48) {
          float2 a = make_float2(0.f, 0.f);
#pragma unroll
          for (int tp = 0; tp < TAPS; ++tp) {
            const float f = __ldg(frac + tp);
            a.x = fmaf(sw2[tp * 48 + wd].x, f, a.x);
            a.y = fmaf(sw2[tp * 48 + wd].y, f, a.y);
          }
          sweff[2 * wd] = a.x;
          sweff[2 * wd + 1] = a.y;
        }
        have_weff = true;
      }
      const int head = item % g.h, b = item / g.h;
      const bf16* zin = in + (int64_t)b * g.in_bs + (int64_t)head * g.in_hs;
      bf16* obase = out + (int64_t)item * Nout * PD;
      for (int base = 0; base < 1 + g.O; base += C::NTOK) {
        __syncthreads();  // ring reads / previous flush finished with the staging tile; w_eff visible
        const int n = min(C::NTOK, 1 + g.O - base);
        for (int i = threadIdx.x; i < n * 48; i += MK_THREADS) {
          const int k = i / 48, ww = i - k * 48;
          const int r = base + k;  // 0 = cls, r >= 1: object token r-1
          const int64_t tok_in = r == 0 ? 0 : L + r;
          const uint32_t v = reinterpret_cast<const uint32_t*>(zin + tok_in * g.in_ts)[ww];
          const float sx = r == 0 ? 1.f : sweff[2 * ww], sy = r == 0 ? 1.f : sweff[2 * ww + 1];
          stg[k * MK_PITCH + ww] = make_float2(lo_f(v) * sx, hi_f(v) * sy);
          if (ww == 0) stg_tok[k] = r == 0 ? 0 : Lo + r;
        }
        __syncthreads();
        mk_ln_flush(stg, stg_tok, aff, n, eps, obase, threadIdx.x, SAVE ? pre + (int64_t)item * Nout * PD : nullptr);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ direct
// One half-warp per output token.  mode 0: all tokens; mode 1: only cls + object tokens (companion of the tiled
// kernel, which writes the patch tokens).
__global__ void __launch_bounds__(256, 4)
pool_ln_direct_kernel(const bf16* __restrict__ in, Geom g, const float* __restrict__ w, const float* __restrict__ frac,
                      const float* __restrict__ gamma, const float* __restrict__ beta, bf16* __restrict__ out, float eps,
                      int mode, bf16* __restrict__ pre) {
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
      // one temporal plane at a time: its 9 x 3 words are requested before the first FMA (zero for padding taps);
      // 27 live registers instead of 81 keep four CTAs per SM resident, which is what hides the load latency
#pragma unroll
      for (int kt = 0; kt < 3; ++kt) {
        const int t = to - 1 + kt;
        if (t < 0 || t >= g.T) continue;
        uint32_t xw[9][3];
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const int hh = ho * g.s - 1 + kh;
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const int ww = wo * g.s - 1 + kw;
            const bool ok = hh >= 0 && hh < g.H && ww >= 0 && ww < g.W;
            const uint32_t* q = reinterpret_cast<const uint32_t*>(
                zin + (ok ? (1 + ((int64_t)t * g.H + hh) * g.W + ww) * g.in_ts : 0));
#pragma unroll
            for (int j = 0; j < 3; ++j) xw[kh * 3 + kw][j] = ok ? __ldg(q + l16 + 16 * j) : 0u;
          }
        }
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const float* wr = sw + (kt * 9 + tap) * PD;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float2 f = *reinterpret_cast<const float2*>(wr + 2 * (l16 + 16 * j));
            v[2 * j] = fmaf(lo_f(xw[tap][j]), f.x, v[2 * j]);
            v[2 * j + 1] = fmaf(hi_f(xw[tap][j]), f.y, v[2 * j + 1]);
          }
        }
      }
    }
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + (((int64_t)b * g.h + head) * Nout + tok) * PD);
    if (pre) {  // training: pre-LayerNorm row for the backward
      uint32_t* pd = reinterpret_cast<uint32_t*>(pre + (((int64_t)b * g.h + head) * Nout + tok) * PD);
#pragma unroll
      for (int j = 0; j < 3; ++j) pd[l16 + 16 * j] = pack2(v[2 * j], v[2 * j + 1]);
    }
    ln_store(v, gm, bt, eps, dst, l16);
  }
}

}  // namespace

#ifdef SVIT_TIMELINE
extern "C" int svit_debug_pool_timeline(void* device_buffer) {
  unsigned long long* p = (unsigned long long*)device_buffer;
  return (int)cudaMemcpyToSymbol(g_pool_dbg, &p, sizeof(p));
}
#endif

namespace {
template <int SV, int SPRV, bool PERSIST, bool SAVE>
int launch_march(const void* in, int64_t in_bs, int64_t in_ts, const Geom& g, const float* conv_w, const float* tap_frac,
                 const float* gamma, const float* beta, void* out, float eps, cudaStream_t st, void* pre) {
  using C = MkCfg<SV, SPRV>;
  // 5-D view of the patch tokens: (head*96 + c, w, h, t, b); the box of a CTA is (96, IW, IH, 1, 1), zero fill outside
  svit_tmap_encode_fn enc = svit_get_tmap_encode();
  if (!enc) return SVIT_ENOTSUP;
  CUtensorMap tm;
  cuuint64_t dims[5] = {(cuuint64_t)g.h * PD, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.T, (cuuint64_t)g.B};
  cuuint64_t strides[4] = {(cuuint64_t)in_ts * 2, (cuuint64_t)g.W * in_ts * 2, (cuuint64_t)g.H * g.W * in_ts * 2, (cuuint64_t)in_bs * 2};
  cuuint32_t box[5] = {PD, (cuuint32_t)C::IW, (cuuint32_t)C::IH, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const bf16* base = (const bf16*)in + in_ts;  // token 0 is the cls token
  if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<bf16*>(base), dims, strides, box, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return SVIT_EINVAL;
  auto kern = pool_ln_march_kernel<SV, SPRV, PERSIST, SAVE>;
  static SvitDevOnce configured;
  if (configured.need()) {
    SVIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    configured.done();
  }
  const int tiles = ((g.Wo + C::TW - 1) / C::TW) * ((g.Ho + C::ROWS - 1) / C::ROWS);
  const int64_t ncols = (int64_t)tiles * g.h * g.B;
  if (ncols > (1ll << 30)) return SVIT_EINVAL;
  int64_t grid = PERSIST ? (int64_t)svit_num_sms() * 2 : ncols;
  if (grid > ncols) grid = ncols;
  kern<<<(unsigned)grid, MK_THREADS, C::SMEM, st>>>(tm, (const bf16*)in, g, conv_w, tap_frac, gamma, beta, (bf16*)out, eps,
                                                    tiles, (int)ncols, (bf16*)pre);
  SVIT_CHECK_LAUNCH();
  return 0;
}
}  // namespace

// bf16 fast path of svit_pool_ln_fwd (pool_ln.cu dispatches here).  Requires 4-byte aligned token slices.
int svit_pool_ln_fwd_bf16(const void* in, int64_t in_bs, int64_t in_ts, int64_t in_hs, const float* conv_w,
                          const float* tap_frac, const float* gamma, const float* beta, void* out, int B, int h, int T,
                          int H, int W, int O, int s, float eps, cudaStream_t st, void* pre) {
  Geom g;
  g.B = B; g.h = h; g.T = T; g.H = H; g.W = W; g.O = O; g.s = s;
  g.Ho = (H - 1) / s + 1; g.Wo = (W - 1) / s + 1;
  g.in_bs = in_bs; g.in_ts = in_ts; g.in_hs = in_hs;
  const int sms = svit_num_sms();
  const bool tma_ok = (s == 1 || s == 2) && (in_ts % 8 == 0) && (in_bs % 8 == 0) && in_hs == PD && (T * H * W) > 0 &&
                      ((reinterpret_cast<uintptr_t>(in) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  if (tma_ok) {
    // Persistent CTAs (static column lists) pay off only where columns are short and many: stride 2 on the large grids
    // (B64 h1 56x56 s2: 112 us against 119 us).  Everywhere else one column per CTA wins: the hardware scheduler
    // balances the uneven tail (3.46 columns per CTA slot at B64 h4 14x14 s1: 95 us against 114 us persistent).
    static const int mode = []() { const char* e = getenv("SVIT_POOL_PERSIST"); return e ? atoi(e) : -1; }();  // -1 auto, 0 never, 1 always
    const bool wide = s == 1 && g.Wo > 7;
    const int64_t cols_s2 = (int64_t)((g.Wo + 6) / 7) * ((g.Ho + 3) / 4) * h * B;
    const bool persist = mode < 0 ? (s == 2 && cols_s2 >= (int64_t)sms * 12) : mode != 0;
#define MARCH(SV, SPRV)                                                                                               \
  (pre ? (persist ? launch_march<SV, SPRV, true, true>(in, in_bs, in_ts, g, conv_w, tap_frac, gamma, beta, out, eps, st, pre)    \
                  : launch_march<SV, SPRV, false, true>(in, in_bs, in_ts, g, conv_w, tap_frac, gamma, beta, out, eps, st, pre)) \
       : (persist ? launch_march<SV, SPRV, true, false>(in, in_bs, in_ts, g, conv_w, tap_frac, gamma, beta, out, eps, st, nullptr) \
                  : launch_march<SV, SPRV, false, false>(in, in_bs, in_ts, g, conv_w, tap_frac, gamma, beta, out, eps, st, nullptr)))
    if (s == 2) return MARCH(2, 2);
    if (!wide) return MARCH(1, 1);
    return MARCH(1, 2);
#undef MARCH
  }
  const int64_t total = (int64_t)B * h * (1 + (int64_t)T * g.Ho * g.Wo + O);
  int64_t blocks = (total + 15) / 16;
  if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
  if (blocks < 1) blocks = 1;
  pool_ln_direct_kernel<<<(unsigned)blocks, 256, 0, st>>>((const bf16*)in, g, conv_w, tap_frac, gamma, beta, (bf16*)out, eps, 0, (bf16*)pre);
  SVIT_CHECK_LAUNCH();
  return 0;
}
