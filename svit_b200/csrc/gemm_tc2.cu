// bf16 GEMM on the 5th-generation tensor cores, TMA-store epilogue (the production path of svit_gemm for bf16
// outputs; gemm_tc.cu keeps the generic-epilogue kernel for fp32 outputs, pre-activation side outputs and
// MN-major A).  Same contract (svit_gemm_args):
//   C[M,N] = residual + sample_scale[row/rps] * ( act(A.op(B) + bias) * gelu'(gelu_pre) )
//
// Persistent, warp-specialised, one CTA per SM (320 or 448 threads):
//   warp 0     TMA producer: ring of {A 128x64, B BNx64} bf16 tiles, 128-byte swizzle, OOB zero fill
//   warp 1     MMA issuer (one lane): tcgen05.mma 128 x BN x 16, fp32 accumulators double-buffered in TMEM
//   warps 2..   epilogue, two or three groups of four warps (one warp per TMEM lane quarter, thread = accumulator row).
//              A group takes every third 32-column box of the tile: tcgen05.ld (32 columns) -> bias / GELU /
//              gelu' / DropPath scale / residual in registers -> bf16 -> 64-byte-swizzled staging slot -> one
//              TMA store per box (cp.async.bulk.tensor, clipped at the tensor edge).  The residual (or gelu_pre)
//              box is TMA-loaded INTO the staging slot two boxes ahead by the group's leader thread and combined
//              in place, so every global access of the epilogue is a bulk tensor copy.
// Ring of SPG staging slots per group: a slot is rewritten only after the store that read it has drained
// (cp.async.bulk.wait_group.read SPG-2 by the leader, published by the group's named barrier).
#include <cstdlib>

#include "tc_common.cuh"
#include "../../include/svit_b200.h"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
// Epilogue shape G = groups of four warps (one warp per TMEM lane quarter): G = 3 with 3 staging slots per group
// for epilogue-heavy problems (GELU, short K), G = 2 with 4 slots (one more operand stage) otherwise.
constexpr int spg_of(int G) { return G == 3 ? 3 : 4; }
constexpr int threads_of(int G) { return 64 + 128 * G; }
constexpr int BOX_N = 32;                 // columns per epilogue box
constexpr int SLOT_BYTES = BM * BOX_N * 2;  // 8 KB
constexpr int PF_DIST = 2;                // aux boxes requested ahead
constexpr int SMEM_LIMIT = 232448;        // 227 KB

enum { AUX_NONE = 0, AUX_RESIDUAL = 1, AUX_GELU_PRE = 2 };

template <int BN, int G, bool PAIR = false>
struct Cfg {
  static constexpr int SPG = spg_of(G);
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * BK * 2;  // a CTA of a pair holds half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STG_BYTES = G * SPG * SLOT_BYTES;
  static constexpr int FIXED = STG_BYTES + 512 + 1024;  // staging + barriers + alignment slack
  static constexpr int STAGES_RAW = (SMEM_LIMIT - FIXED) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int STG_OFF = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFF = STG_OFF + STG_BYTES;
  static constexpr int TOTAL = BAR_OFF + 512 + 1024;
  static constexpr uint32_t TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
};

struct Epi {
  const float* bias;
  const float* sample_scale;
  int64_t rps;
  int act;
  int aux;  // AUX_*
  int64_t rows_in, rows_out, row_off;
  const float* ln_stats;   // folded LayerNorm: (mean, rstd) per row of A, or NULL
  const float* ln_colsum;  // ... and sum_k B[n, k]
  unsigned long long* dbg;  // optional timeline buffer (CTA 0 only): [role][event] = (tag, clock64)
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7): Phi(x) and x * phi(x) for gelu'(x) = Phi(x) + x phi(x).
__device__ __forceinline__ float gelu_grad(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float ex = exp2f(-1.4426950408889634f * z * z);
  const float erfc_half = 0.5f * poly * t * ex;
  const float cdf = x >= 0.f ? 1.0f - erfc_half : erfc_half;
  return fmaf(x * 0.3989422804014327f, ex, cdf);
}

using tc::add2;
using tc::fma2;
using tc::mul2;

// x * Phi(x) on two elements, Phi through a tanh form fitted to the exact erf GELU (max abs deviation 2.5e-5 on
// [-9, 9], x^2 clamped beyond) and MUFU.TANH; the error is an order of magnitude below the bf16 rounding of the
// result.  Per pair: 5 packed FP32 ops + 2 FMNMX + 2 MUFU.
__device__ __forceinline__ float2 gelu_fwd_fast2(const float2 x) {
  float2 x2 = mul2(x, x);
  x2.x = fminf(x2.x, 81.0f);
  x2.y = fminf(x2.y, 81.0f);
  const float2 p = fma2(x2, fma2(x2, make_float2(-3.51516789e-04f, -3.51516789e-04f), make_float2(3.70056460e-02f, 3.70056460e-02f)),
                        make_float2(7.97507884e-01f, 7.97507884e-01f));
  const float2 u = mul2(x, p);
  float2 th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th.x) : "f"(u.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(th.y) : "f"(u.y));
  const float2 hx = mul2(x, make_float2(0.5f, 0.5f));
  return fma2(hx, th, hx);
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)map),
               "r"(tc::smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Static schedule of one epilogue group's boxes: tiles blockIdx.x, +gridDim.x, ...; inside a tile the boxes
// c = g, g+2, ... that start left of N.
struct BoxIter {
  int t, num_tiles, n_tiles, N;
  int c, g, BN, G, stride;
  __device__ BoxIter(int t0, int stride_, int num_tiles_, int n_tiles_, int N_, int g_, int BN_, int G_)
      : t(t0), num_tiles(num_tiles_), n_tiles(n_tiles_), N(N_), c(g_), g(g_), BN(BN_), G(G_), stride(stride_) {
    settle();
  }
  __device__ int nbox() const {
    const int n0 = (t % n_tiles) * BN;
    const int rem = N - n0 < BN ? N - n0 : BN;
    return (rem + BOX_N - 1) / BOX_N;
  }
  __device__ void settle() {
    while (t < num_tiles && c >= nbox()) {
      t += stride;
      c = g;
    }
  }
  __device__ bool valid() const { return t < num_tiles; }
  __device__ void next() {
    c += G;
    settle();
  }
};

// timeline probe (build with SVIT_NVCC_EXTRA=-DSVIT_TIMELINE; tools/gemm_timeline.py): role 0 producer, 1 MMA,
// 2 epilogue group-0 leader lane; 4096 events per role
#ifdef SVIT_TIMELINE
#define TL(role, tag)                                                            \
  do {                                                                           \
    if (e.dbg && blockIdx.x == 0 && tl_n < 4096) {                               \
      e.dbg[((role) * 4096 + tl_n) * 2] = (unsigned long long)(tag);             \
      e.dbg[((role) * 4096 + tl_n) * 2 + 1] = (unsigned long long)clock64();     \
      ++tl_n;                                                                    \
    }                                                                            \
  } while (0)
#else
#define TL(role, tag) do { (void)tl_n; } while (0)
#endif

__device__ __forceinline__ int remap_row(const Epi& e, int m) {  // all row counts are < 2^31 (svit_gemm_tc_supported)
  if (e.rows_in <= 0) return m;
  const int ri = (int)e.rows_in, q = m / ri;
  return q * (int)e.rows_out + (int)e.row_off + (m - q * ri);
}

// PAIR: clusters of two CTAs compute 256 x BN tiles with tcgen05.mma.cta_group::2 -- each CTA loads its 128 rows of A
// and HALF of the B tile, the leader CTA issues the MMAs for both, completion is multicast to both CTAs' barriers.
template <int BN, bool B_MN, int G, bool PAIR, bool LN>
__global__ void __launch_bounds__(threads_of(G), 1)
gemm_tc_tma_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                   const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_x, int M,
                   int N, int K, Epi e) {
  using L = Cfg<BN, G, PAIR>;
  static_assert(!(PAIR && B_MN), "pair mode supports K-major B only");
  constexpr int STAGES = L::STAGES;
  constexpr int TM = PAIR ? 2 * BM : BM;  // tile rows of the scheduling unit (CTA or CTA pair)
  constexpr int SPG = L::SPG;
  constexpr int EPI_WARPS = 4 * G;
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment by pointer offset (not an integer round trip) so accesses stay in the shared state space
  unsigned char* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* stg_base = smem + L::STG_OFF;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* aux_full = tmem_empty + 2;  // [G * SPG]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(aux_full + G * SPG);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (M + TM - 1) / TM, n_tiles = (N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (K + BK - 1) / BK;
  const uint32_t crank = PAIR ? tc::cluster_ctarank() : 0u;   // 0 = leader of the pair
  const int unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int unit_stride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_a);
    tc::prefetch_tmap(&tmap_b);
    tc::prefetch_tmap(&tmap_c);
    if (e.aux) tc::prefetch_tmap(&tmap_x);
    for (int i = 0; i < STAGES; ++i) {
      tc::mbar_init(&full_bar[i], 1);
      tc::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&tmem_full[i], 1);
      tc::mbar_init(&tmem_empty[i], PAIR ? 2 * EPI_WARPS : EPI_WARPS);  // pair: both CTAs' epilogues report to the leader
    }
    for (int i = 0; i < G * SPG; ++i) tc::mbar_init(&aux_full[i], 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) tc::tmem_alloc_pair(tmem_ptr, L::TMEM_COLS);
    else tc::tmem_alloc(tmem_ptr, L::TMEM_COLS);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (PAIR) tc::cluster_sync();  // the peer's barriers are initialised before anything signals them
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (tc::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int tl_n = 0;
      for (int t = unit0; t < num_tiles; t += unit_stride) {
        const int m0 = (t / n_tiles) * TM + (int)crank * BM, n0 = (t % n_tiles) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          tc::mbar_wait(&empty_bar[stage], phase ^ 1);
          TL(0, kb);
          unsigned char* sa = smem + stage * L::STAGE_BYTES;
          unsigned char* sb = sa + L::A_BYTES;
          if (PAIR) {
            // both CTAs' bytes are accounted on the leader's barrier, which the leader arms for the pair
            const uint32_t lbar = tc::mapa_shared(tc::smem_u32(&full_bar[stage]), 0);
            if (crank == 0) tc::mbar_arrive_expect_tx(&full_bar[stage], 2 * L::STAGE_BYTES);
            tc::tma_load_2d_pair(sa, &tmap_a, lbar, kb * BK, m0);
            tc::tma_load_2d_pair(sb, &tmap_b, lbar, kb * BK, n0 + (int)crank * (BN / 2));
          } else {
            tc::mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
            tc::tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, m0);
            if (!B_MN) {
              tc::tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BK, n0);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j) tc::tma_load_2d(sb + j * 8192, &tmap_b, &full_bar[stage], n0 + 64 * j, kb * BK);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (crank == 0 && tc::elect_one()) {
      constexpr uint32_t idesc = tc::idesc_bf16(TM, BN, 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      int tl_n = 0;
      for (int t = unit0; t < num_tiles; t += unit_stride, ++it) {
        const int as = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        tc::mbar_wait_hot(&tmem_empty[as], acc_phase ^ 1);
        tc::fence_after_sync();
        TL(1, 1000);
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          tc::mbar_wait_hot(&full_bar[stage], phase);
          tc::fence_after_sync();
          TL(1, kb);
          const uint32_t sa = tc::smem_u32(smem + stage * L::STAGE_BYTES);
          const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = tc::smem_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? tc::smem_desc_sw128(sb + k * 2048, 8192, 1024) : tc::smem_desc_sw128(sb + k * 32, 16, 1024);
            if (PAIR) tc::umma_bf16_ss_pair(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            else tc::umma_bf16_ss(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if (PAIR) {
            tc::umma_commit_pair(&empty_bar[stage]);
            if (kb == num_kb - 1) tc::umma_commit_pair(&tmem_full[as]);
          } else {
            tc::umma_commit(&empty_bar[stage]);
            if (kb == num_kb - 1) tc::umma_commit(&tmem_full[as]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int g = (warp - 2) >> 2;     // epilogue group: boxes c with c % G == g
    const int r = q * 32 + lane;       // accumulator row inside the tile
    const bool leader = (warp - 2) == g * 4 && lane == 0;
    unsigned char* my_slots = stg_base + g * SPG * SLOT_BYTES;
    uint64_t* my_aux = aux_full + g * SPG;
    const uint32_t sw = (uint32_t)((r >> 1) & 3);
    const int aux = e.aux;

    // leader: request the first PF_DIST aux boxes
    BoxIter pf(unit0, unit_stride, num_tiles, n_tiles, N, g, BN, G);
    auto release_tmem = [&](int as) {  // one arrival per epilogue warp on the (leader's) tmem_empty barrier
      if (PAIR) tc::mbar_arrive_cluster(tc::mapa_shared(tc::smem_u32(&tmem_empty[as]), 0));
      else tc::mbar_arrive(&tmem_empty[as]);
    };
    uint32_t pf_cnt = 0;
    auto request_aux = [&]() {
      if (!pf.valid()) return;
      const int m0 = (pf.t / n_tiles) * TM + (int)crank * BM, n0 = (pf.t % n_tiles) * BN;
      const int row0 = aux == AUX_RESIDUAL ? remap_row(e, m0) : m0;
      const int slot = pf_cnt % SPG;
      tc::mbar_arrive_expect_tx(&my_aux[slot], SLOT_BYTES);
      tc::tma_load_2d(my_slots + slot * SLOT_BYTES, &tmap_x, &my_aux[slot], n0 + pf.c * BOX_N, row0);
      ++pf_cnt;
      pf.next();
    };
    if (leader && aux) {
      for (int i = 0; i < PF_DIST; ++i) request_aux();
    }

    // folded LayerNorm: (mean, rstd) of this thread's row, fetched ONE TILE AHEAD (the load misses L1/L2: ~1 us that
    // must not sit between the accumulator becoming ready and the first box)
    auto load_stats = [&](int t) {
      const int m = (t / n_tiles) * TM + (int)crank * BM + r;
      return __ldg(reinterpret_cast<const float2*>(e.ln_stats) + (m < M ? m : M - 1));
    };
    float2 st_next = make_float2(0.f, 1.f);
    if (LN && unit0 < num_tiles) st_next = load_stats(unit0);

    uint32_t cnt = 0;
    int it = 0;
    int tl_n = leader ? 0 : 4096;
    if (g != 0) tl_n = 4096;
    for (int t = unit0; t < num_tiles; t += unit_stride, ++it) {
      const int as = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m0 = (t / n_tiles) * TM + (int)crank * BM, n0 = (t % n_tiles) * BN;
      const int rem = N - n0 < BN ? N - n0 : BN;
      const int nbox = (rem + BOX_N - 1) / BOX_N;
      const int orow0 = remap_row(e, m0);
      float sc = 1.f;
      if (e.sample_scale) {
        const int m = m0 + r < M ? m0 + r : M - 1;
        sc = e.sample_scale[m / (int)e.rps];
      }
      float2 ln_a = make_float2(1.f, 1.f), ln_b = make_float2(0.f, 0.f);  // rstd, -rstd * mean of this thread's row
      if (LN) {
        const float2 st = st_next;
        if (t + unit_stride < num_tiles) st_next = load_stats(t + unit_stride);
        ln_a = make_float2(st.y, st.y);
        ln_b = make_float2(-st.y * st.x, -st.y * st.x);
      }
      int last_c = -1;
      for (int c = g; c < nbox; c += G) last_c = c;
      TL(2, 2000);
      tc::mbar_wait_hot(&tmem_full[as], acc_phase);
      tc::fence_after_sync();
      TL(2, 2001);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
      if (last_c < 0) {  // no box for this group in this tile (narrow tail): keep the arrival count uniform
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) release_tmem(as);
      }
#pragma unroll 1
      for (int c = g; c < nbox; c += G) {
        const int slot = cnt % SPG;
        unsigned char* sbase = my_slots + slot * SLOT_BYTES;
        float v[BOX_N];
        tc::tmem_ld32(taddr + c * BOX_N, v);
        const int ncol = n0 + c * BOX_N;
        // bias (L1-resident broadcast loads) while the TMEM load is in flight
        float bv[BOX_N];
        if (!LN && e.bias) {
#pragma unroll
          for (int j = 0; j < BOX_N; j += 4) {
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ncol + j < N) b4 = __ldg(reinterpret_cast<const float4*>(e.bias + ncol + j));
            bv[j] = b4.x; bv[j + 1] = b4.y; bv[j + 2] = b4.z; bv[j + 3] = b4.w;
          }
        }
        tc::tmem_ld_wait();
        TL(2, 100 + c);
        if (c == last_c) {  // this warp's rows of the accumulator are in registers: release the TMEM buffer
          tc::fence_before_sync();
          __syncwarp();
          if (lane == 0) release_tmem(as);
        }
        float2* v2 = reinterpret_cast<float2*>(v);
        if (LN) {
          // LayerNorm(x) W^T + b = rstd * (x W'^T) + (-rstd * mean) * colsum(W') + b'; e.ln_colsum is the interleaved
          // table [N / 2][4] = (colsum[2i], colsum[2i + 1], b'[2i], b'[2i + 1]): one 16-byte load per column pair
#pragma unroll
          for (int j = 0; j < BOX_N / 2; ++j) {
            float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ncol + 2 * j < N) t4 = __ldg(reinterpret_cast<const float4*>(e.ln_colsum) + ((ncol >> 1) + j));
            v2[j] = fma2(v2[j], ln_a, fma2(ln_b, make_float2(t4.x, t4.y), make_float2(t4.z, t4.w)));
          }
        }
        if (!LN && e.bias) {
          const float2* b2 = reinterpret_cast<const float2*>(bv);
#pragma unroll
          for (int j = 0; j < BOX_N / 2; ++j) v2[j] = add2(v2[j], b2[j]);
        }
        if (e.act == 1) {
#pragma unroll
          for (int j = 0; j < BOX_N / 2; ++j) v2[j] = gelu_fwd_fast2(v2[j]);
        }
        uint4 ax[4];
        if (aux) {
          tc::mbar_wait_hot(&my_aux[slot], (cnt / SPG) & 1);
#pragma unroll
          for (int k = 0; k < 4; ++k) ax[k] = *reinterpret_cast<const uint4*>(sbase + r * 64 + ((k ^ sw) << 4));
        }
        if (aux == AUX_GELU_PRE) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const __nv_bfloat162* gp = reinterpret_cast<const __nv_bfloat162*>(&ax[k]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 f = __bfloat1622float2(gp[i]);
              v[k * 8 + 2 * i] *= gelu_grad(f.x);
              v[k * 8 + 2 * i + 1] *= gelu_grad(f.y);
            }
          }
        }
        if (e.sample_scale) {
          const float2 sc2 = make_float2(sc, sc);
#pragma unroll
          for (int j = 0; j < BOX_N / 2; ++j) v2[j] = mul2(v2[j], sc2);
        }
        if (aux == AUX_RESIDUAL) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const __nv_bfloat162* gp = reinterpret_cast<const __nv_bfloat162*>(&ax[k]);
#pragma unroll
            for (int i = 0; i < 4; ++i) v2[k * 4 + i] = add2(v2[k * 4 + i], __bfloat1622float2(gp[i]));
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint4 o = {pack_bf16(v[k * 8], v[k * 8 + 1]), pack_bf16(v[k * 8 + 2], v[k * 8 + 3]),
                           pack_bf16(v[k * 8 + 4], v[k * 8 + 5]), pack_bf16(v[k * 8 + 6], v[k * 8 + 7])};
          *reinterpret_cast<uint4*>(sbase + r * 64 + ((k ^ sw) << 4)) = o;
        }
        TL(2, 200 + c);
        tc::fence_proxy_async();
        named_bar_sync(1 + g, 128);
        TL(2, 300 + c);
        if (leader) {
          tma_store_2d(&tmap_c, sbase, ncol, orow0);
          bulk_commit();
          bulk_wait_read<SPG - 2>();  // the stores of boxes <= cnt - (SPG - 2) have drained: their slots may be refilled
          if (aux) request_aux();
        }
        TL(2, 400 + c);
        ++cnt;
      }
    }
    if (leader) bulk_wait_all();
  }
  tc::fence_before_sync();
  __syncthreads();
  if (PAIR) tc::cluster_sync();  // no CTA leaves (or frees TMEM) while its peer may still signal it
  if (warp == 1) {
    tc::fence_after_sync();
    if (PAIR) tc::tmem_dealloc_pair(tmem_base, L::TMEM_COLS);
    else tc::tmem_dealloc(tmem_base, L::TMEM_COLS);
  }
}

// 2-D bf16 tensor [rows, cols] (row pitch ld), box = [BM rows, 32 cols], 64-byte swizzle (epilogue boxes).
int make_box_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld) {
  svit_tmap_encode_fn enc = svit_get_tmap_encode();
  if (!enc) return SVIT_ENOTSUP;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {BOX_N, BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : SVIT_EINVAL;
}

// SVIT_GEMM_PAIR=0 switches the cta_group::2 path off (diagnostics)
bool svit_gemm_pair_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SVIT_GEMM_PAIR");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

unsigned long long* g_timeline = nullptr;
unsigned long long* svit_gemm_timeline_buffer() { return g_timeline; }

template <int BN, bool B_MN, int G, bool PAIR = false, bool LN = false>
int launch_g(const svit_gemm_args* a, cudaStream_t st) {
  using L = Cfg<BN, G, PAIR>;
  static_assert(L::STAGES >= 2, "pipeline too shallow");
  static_assert(L::TOTAL <= SMEM_LIMIT, "shared memory budget");
  CUtensorMap ta, tb, tcm, tx;
  int rc;
  if ((rc = svit_make_tmap_2d(&ta, a->A, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda, BM))) return rc;
  if (!B_MN) rc = svit_make_tmap_2d(&tb, a->B, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldb, PAIR ? BN / 2 : BN);
  else rc = svit_make_tmap_2d(&tb, a->B, (uint64_t)a->K, (uint64_t)a->N, (uint64_t)a->ldb, BK);
  if (rc) return rc;
  const uint64_t out_rows = a->rows_in > 0 ? (uint64_t)((a->M + a->rows_in - 1) / a->rows_in * a->rows_out) : (uint64_t)a->M;
  if ((rc = make_box_map(&tcm, a->C, out_rows, (uint64_t)a->N, (uint64_t)a->ldc))) return rc;
  Epi e;
  e.bias = a->bias;
  e.sample_scale = a->sample_scale; e.rps = a->rows_per_sample;
  e.act = a->act;
  e.aux = a->residual ? AUX_RESIDUAL : (a->gelu_pre ? AUX_GELU_PRE : AUX_NONE);
  e.rows_in = a->rows_in; e.rows_out = a->rows_out; e.row_off = a->row_off;
  e.ln_stats = a->ln_stats; e.ln_colsum = a->ln_colsum;
  e.dbg = svit_gemm_timeline_buffer();
  tx = tcm;
  if (e.aux == AUX_RESIDUAL) rc = make_box_map(&tx, a->residual, out_rows, (uint64_t)a->N, (uint64_t)a->ldr);
  else if (e.aux == AUX_GELU_PRE) rc = make_box_map(&tx, a->gelu_pre, (uint64_t)a->M, (uint64_t)a->N, (uint64_t)a->ldg);
  if (rc) return rc;
  auto kern = gemm_tc_tma_kernel<BN, B_MN, G, PAIR, LN>;
  static SvitDevOnce configured;
  if (configured.need()) {
    SVIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured.done();
  }
  constexpr int TM = PAIR ? 2 * BM : BM;
  const int64_t tiles = ((a->M + TM - 1) / TM) * ((a->N + BN - 1) / BN);
  if (!PAIR) {
    const int grid = (int)(tiles < svit_num_sms() ? tiles : svit_num_sms());
    kern<<<grid, threads_of(G), L::TOTAL, st>>>(ta, tb, tcm, tx, (int)a->M, (int)a->N, (int)a->K, e);
    SVIT_CHECK_LAUNCH();
    return 0;
  }
  // clusters of two CTAs (one SM pair each)
  int pairs = svit_num_sms() / 2;
  if (tiles < pairs) pairs = (int)tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * pairs));
  cfg.blockDim = dim3((unsigned)threads_of(G));
  cfg.dynamicSmemBytes = L::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SVIT_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, tcm, tx, (int)a->M, (int)a->N, (int)a->K, e));
  return 0;
}

template <int BN, bool B_MN>
int launch(const svit_gemm_args* a, cudaStream_t st) {
  // three epilogue groups when the epilogue dominates: GELU / gelu' math, or few K-steps per tile
  // CTA pairs (256 x BN tiles, cta_group::2) when the reduction is long enough for operand delivery to dominate
  // (measured on B200: +3..12 % for K >= 768; for K = 384 the extra epilogue coupling of the pair costs more than the
  // halved B traffic saves)
  constexpr bool CAN_PAIR = !B_MN && (BN == 256 || BN == 192 || BN == 128);
  if constexpr (!B_MN) {
    if (a->ln_stats) {  // folded LayerNorm (K = the block width: 96 .. 768)
      if constexpr (CAN_PAIR) {
        if (svit_gemm_pair_enabled() && a->K >= 768 && a->M >= 2048) {
          if (a->act == 1) return launch_g<BN, B_MN, 3, true, true>(a, st);
          return launch_g<BN, B_MN, 2, true, true>(a, st);
        }
      }
      if (a->act == 1 || a->K <= 256) return launch_g<BN, B_MN, 3, false, true>(a, st);
      return launch_g<BN, B_MN, 2, false, true>(a, st);
    }
  }
  if constexpr (CAN_PAIR) {
    if (svit_gemm_pair_enabled() && a->K >= 768 && a->M >= 2048) {
      if (a->act == 1 || a->gelu_pre) return launch_g<BN, B_MN, 3, true>(a, st);
      return launch_g<BN, B_MN, 2, true>(a, st);
    }
  }
  if (a->act == 1 || a->gelu_pre || a->K <= 256) return launch_g<BN, B_MN, 3>(a, st);
  return launch_g<BN, B_MN, 2>(a, st);
}

// tile width: the widest of {256, 192, 128, 96, 64} that divides N, else the one wasting the fewest MMA columns
int pick_bn(int64_t N, bool b_mn) {
  const int cand[5] = {256, 192, 128, 96, 64};
  for (int i = 0; i < 5; ++i) {
    if (b_mn && cand[i] % 64) continue;
    if (N % cand[i] == 0) return cand[i];
  }
  int best = 64;
  int64_t best_cols = -1;
  for (int i = 0; i < 5; ++i) {
    if (b_mn && cand[i] % 64) continue;
    const int64_t cols = (N + cand[i] - 1) / cand[i] * cand[i];
    if (best_cols < 0 || cols < best_cols) { best_cols = cols; best = cand[i]; }
  }
  return best;
}

template <bool B_MN>
int dispatch(const svit_gemm_args* a, cudaStream_t st) {
  switch (pick_bn(a->N, B_MN)) {
    case 256: return launch<256, B_MN>(a, st);
    case 192: return launch<192, B_MN>(a, st);
    case 128: return launch<128, B_MN>(a, st);
    case 96: if (!B_MN) return launch<96, false>(a, st); return launch<128, B_MN>(a, st);
    default: return launch<64, B_MN>(a, st);
  }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

// Preconditions beyond svit_gemm_tc_supported(): bf16 output, K-major A, at most one auxiliary input, no
// pre-activation side output, output row remap only with whole tiles per segment.
int svit_gemm_tc_tma_supported(const svit_gemm_args* a) {
  if (a->out_dtype != SVIT_BF16 || a->transA) return 0;
  if (a->pre_out) return 0;
  if (a->residual && a->gelu_pre) return 0;
  if (a->rows_in > 0 && (a->rows_in % BM || a->gelu_pre)) return 0;
  if (a->ldc % 8 || !aligned16(a->C)) return 0;
  if (a->residual && (a->ldr % 8 || !aligned16(a->residual))) return 0;
  if (a->gelu_pre && (a->ldg % 8 || !aligned16(a->gelu_pre))) return 0;
  if (a->bias && !aligned16(a->bias)) return 0;
  if (a->ln_stats && (!a->ln_colsum || !aligned16(a->ln_colsum) || (reinterpret_cast<uintptr_t>(a->ln_stats) & 7) || a->N % 2 ||
                      a->transB == 0 || a->gelu_pre || a->residual || a->sample_scale))
    return 0;
  if (a->M < BM) return 0;  // tiny problems (heads): the generic kernel is fine
  return 1;
}

// Diagnostic hook (not part of the public header): device buffer of 3 * 4096 * 2 uint64 that CTA 0 of the next
// launches fills with (tag, clock64) events; pass NULL to switch the probe off.
extern "C" int svit_debug_gemm_timeline(void* device_buffer) {
  g_timeline = (unsigned long long*)device_buffer;
  return 0;
}

int svit_gemm_tc_tma(const svit_gemm_args* a, cudaStream_t st) {
  if (a->transB == 0) return dispatch<true>(a, st);
  return dispatch<false>(a, st);
}
