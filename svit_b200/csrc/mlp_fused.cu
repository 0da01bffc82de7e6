// Fused MLP of a MultiScaleBlock (inference): out = residual + fc2(gelu(fc1(LN(x))))     slowfast/models/common.py:27-34,
// slowfast/models/attention.py:566-570.  The hidden activation [M, 4C] never reaches HBM: per 128-row tile the hidden
// units are produced 64 at a time into TMEM, activated by the epilogue warps, written to shared memory as the bf16 A
// operand of the second product and consumed from there.  For the wide, short stages (C = 96 with 1.6 M rows, C = 192 with
// 0.4 M rows at batch 64) fc1 / fc2 as two GEMMs are HBM-bound on exactly that tensor (write 4C, read 4C per row against
// C in, C out).
//
// Persistent, one CTA per SM.  Roles:
//   producer warp(s)  TMA: x tile [128, C] (resident for the tile); C = 96: both weight matrices once per CTA (144 KB
//                     resident), C = 192: weight chunks through two rings
//   MMA issuer        one lane:  acc1[g % 4] = x W1_c^T  (128 x 64 x C),  acc2 += h_c W2_c^T  (128 x NOUT x 64); the first
//                     product runs up to three chunks ahead of the second
//   2 x 8 epilogue warps ("teams", two warps per TMEM lane quarter, 32 of a chunk's 64 hidden units each):
//                     tcgen05.ld -> + b1 -> GELU -> bf16 -> 128-byte-swizzled K-major tile in shared memory -> mbarrier;
//                     per tile the LayerNorm prologue (C = 96) and the output (acc2 + b2 + residual -> TMA store)
// TMEM: four first-product accumulators (4 x 64 columns) | acc2 (NOUT columns).
#include "tc_common.cuh"
#include "../../include/svit_b200.h"

namespace {

constexpr int BM = 128;
constexpr int HC = 64;                       // hidden units per chunk
constexpr int BOX_N = 32;                    // output columns per staging slot
constexpr int SLOT_BYTES = BM * BOX_N * 2;   // 8 KB
constexpr int SMEM_LIMIT = 232448;

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
using tc::add2;
using tc::fma2;
using tc::mul2;

// x * Phi(x) on two elements: the tanh form fitted to the exact erf GELU that the GEMM epilogue uses (gemm_tc2.cu;
// max abs deviation 2.5e-5 on [-9, 9], an order of magnitude below the bf16 rounding of the result)
__device__ __forceinline__ float2 gelu2(const float2 x) {
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

__device__ __forceinline__ uint64_t smem_desc_sw64(uint32_t smem_addr) {  // K-major, 64-byte rows, 8-row groups 512 B apart
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(16 >> 4) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)map),
               "r"(tc::smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"((uint64_t)map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// C = 96, H = 384, NOUT = 96 (the first MultiScaleBlock: 25 153 tokens per clip): BOTH weight matrices stay resident in
// shared memory (144 KB, loaded once per CTA), so a tile moves only its rows: x in, out back.  Optionally the block's
// LayerNorm runs as the tile's prologue (x arrives un-normalised, the epilogue threads normalise it in place in shared
// memory and keep their part of the raw row in registers as the residual): out = x + fc2(gelu(fc1(LayerNorm(x))))
// (attention.py:566-570) with one read and one write of the activations.
namespace r96 {
constexpr int C = 96, H = 384, NOUT = 96, NC = H / HC;                    // 6 chunks
constexpr int XS_BYTES = BM * 128 + BM * 64;                               // 24 KB: columns 0..63 (SW128) | 64..95 (SW64)
constexpr int W1C_BYTES = HC * 128 + HC * 64;                              // 12 KB per chunk
constexpr int W2C_BYTES = NOUT * 128;                                      // 12 KB per chunk
constexpr int OFF_XS = 0;
constexpr int OFF_W1 = OFF_XS + XS_BYTES;
constexpr int OFF_W2 = OFF_W1 + NC * W1C_BYTES;
constexpr int OFF_HS = OFF_W2 + NC * W2C_BYTES;                            // 2 x 16 KB hidden tiles (one per team)
constexpr int OFF_SLOT = OFF_HS + 2 * BM * 128;                            // 2 x 8 KB output staging (one per column group)
constexpr int OFF_XCH = OFF_SLOT + 2 * SLOT_BYTES;                         // [2][2][128] fp32 LayerNorm partial sums
constexpr int OFF_BAR = OFF_XCH + 4 * BM * 4;
constexpr int TOTAL = OFF_BAR + 256 + 1024;
constexpr int NA1 = 4;                                                     // first-product accumulators (two per team)
constexpr int LOOK = NA1 - 1;                                              // chunks the first product runs ahead of the second
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t COL_ACC2 = NA1 * HC;
constexpr int TEAM_WARPS = 8;
constexpr int NTHR = 64 + 2 * TEAM_WARPS * 32;                             // 576
static_assert(TOTAL <= SMEM_LIMIT, "shared memory budget");
enum { B_W_FULL = 0, B_XS_FULL, B_XS_READY, B_XS_EMPTY, B_ACC1_FULL0, B_ACC1_EMPTY0 = B_ACC1_FULL0 + NA1,
       B_HS_FULL0 = B_ACC1_EMPTY0 + NA1, B_HS_FULL1, B_HS_EMPTY0, B_HS_EMPTY1, B_ACC2_FULL, B_ACC2_EMPTY, NUM_BARS };
}  // namespace r96

// timeline probe (build with -DSVIT_TIMELINE; tools/mlp_timeline.py): role 0 producer, 1 MMA, 2 first warp of epilogue
// team TL_TEAM
#ifndef TL_TEAM
#define TL_TEAM 0
#endif
#ifdef SVIT_TIMELINE
#define TLR(role, tag)                                                           \
  do {                                                                           \
    if (p.dbg && blockIdx.x == 0 && tl_n < 4096) {                               \
      p.dbg[((role) * 4096 + tl_n) * 2] = (unsigned long long)(tag);             \
      p.dbg[((role) * 4096 + tl_n) * 2 + 1] = (unsigned long long)clock64();     \
      ++tl_n;                                                                    \
    }                                                                            \
  } while (0)
#else
#define TLR(role, tag) do { (void)tl_n; } while (0)
#endif
unsigned long long* g_mlp_timeline = nullptr;

struct ParamsR {
  unsigned long long* dbg;
  const float* b1;
  const float* b2;
  const float* gamma;     // LayerNorm prologue (NULL: x is used as it is)
  const float* beta;
  const bf16* residual;   // residual rows [M, 96] (x itself in the LayerNorm mode), or NULL
  float eps;
  int M;
};

// Two epilogue teams of eight warps (two warps per TMEM lane quarter, 32 hidden units of a chunk each):
//   team 0: LayerNorm prologue of the tile, hidden chunks 0, 2, 4 (accumulator / hidden buffer 0)
//   team 1: hidden chunks 1, 3, 5 (buffer 1), output epilogue of the tile
// Every phase of such an epilogue is a chain of latencies (mbarrier -> tcgen05.ld -> MUFU -> st.shared -> proxy fence ->
// mbarrier; measured ~1.5 k cycles per chunk and ~2.5 k each for the prologue and the output with ONE team, the tensor
// pipe one third busy), so the two teams run the chains of neighbouring chunks -- and of neighbouring tiles -- side by side.
__global__ void __launch_bounds__(r96::NTHR, 1)
mlp_resident96_kernel(const __grid_constant__ CUtensorMap tmap_x64, const __grid_constant__ CUtensorMap tmap_x32,
                      const __grid_constant__ CUtensorMap tmap_w1_64, const __grid_constant__ CUtensorMap tmap_w1_32,
                      const __grid_constant__ CUtensorMap tmap_w2, const __grid_constant__ CUtensorMap tmap_out, ParamsR p) {
  using namespace r96;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + NUM_BARS);
  float* xch = reinterpret_cast<float*>(smem + OFF_XCH);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = (p.M + BM - 1) / BM;
  const bool ln = p.gamma != nullptr;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_x64); tc::prefetch_tmap(&tmap_x32); tc::prefetch_tmap(&tmap_w1_64); tc::prefetch_tmap(&tmap_w1_32);
    tc::prefetch_tmap(&tmap_w2); tc::prefetch_tmap(&tmap_out);
    for (int i = 0; i < NUM_BARS; ++i) {
      const bool per_warp = i == B_XS_READY || (i >= B_ACC1_EMPTY0 && i < B_ACC1_EMPTY0 + NA1) || i == B_HS_FULL0 ||
                            i == B_HS_FULL1 || i == B_ACC2_EMPTY;
      tc::mbar_init(&bars[i], per_warp ? TEAM_WARPS : 1);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_ptr, TMEM_COLS);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (tc::elect_one()) {
      tc::mbar_arrive_expect_tx(&bars[B_W_FULL], NC * (W1C_BYTES + W2C_BYTES));
      for (int c = 0; c < NC; ++c) {
        unsigned char* w1 = smem + OFF_W1 + c * W1C_BYTES;
        tc::tma_load_2d(w1, &tmap_w1_64, &bars[B_W_FULL], 0, c * HC);
        tc::tma_load_2d(w1 + HC * 128, &tmap_w1_32, &bars[B_W_FULL], 64, c * HC);
        tc::tma_load_2d(smem + OFF_W2 + c * W2C_BYTES, &tmap_w2, &bars[B_W_FULL], c * HC, 0);
      }
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int m0 = t * BM;
        tc::mbar_wait(&bars[B_XS_EMPTY], (it & 1) ^ 1);
        tc::mbar_arrive_expect_tx(&bars[B_XS_FULL], XS_BYTES);
        tc::tma_load_2d(smem + OFF_XS, &tmap_x64, &bars[B_XS_FULL], 0, m0);
        tc::tma_load_2d(smem + OFF_XS + BM * 128, &tmap_x32, &bars[B_XS_FULL], 64, m0);
        if (t + (int)gridDim.x < num_tiles) {  // the next tile's rows: into L2 now, into shared memory when the buffer frees up
          tma_prefetch_l2_2d(&tmap_x64, 0, m0 + (int)gridDim.x * BM);
          tma_prefetch_l2_2d(&tmap_x32, 64, m0 + (int)gridDim.x * BM);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (tc::elect_one()) {
      constexpr uint32_t idesc1 = tc::idesc_bf16(BM, HC, 0, 0);
      constexpr uint32_t idesc2 = tc::idesc_bf16(BM, NOUT, 0, 0);
      const uint32_t xs = tc::smem_u32(smem + OFF_XS);
      tc::mbar_wait(&bars[B_W_FULL], 0);
      uint32_t g1 = 0, g2 = 0;
      int it = 0;
      int tl_n = 0;
      auto mma1 = [&](int c) {  // acc1[g1 % NA1] = x W1_c^T  (the six chunks of a tile keep their parity: g1 & 1 == c & 1)
        const int b = g1 & (NA1 - 1);
        tc::mbar_wait(&bars[B_ACC1_EMPTY0 + b], ((g1 / NA1) & 1) ^ 1);
        tc::fence_after_sync();
        TLR(1, 100 + c);
        const uint32_t w1 = tc::smem_u32(smem + OFF_W1 + c * W1C_BYTES);
        const uint32_t d = tmem_base + (uint32_t)(b * HC);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc::umma_bf16_ss(d, tc::smem_desc_sw128(xs + k * 32, 16, 1024), tc::smem_desc_sw128(w1 + k * 32, 16, 1024), idesc1,
                           k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 2; ++k)
          tc::umma_bf16_ss(d, smem_desc_sw64(xs + BM * 128 + k * 32), smem_desc_sw64(w1 + HC * 128 + k * 32), idesc1, 1u);
        tc::umma_commit(&bars[B_ACC1_FULL0 + b]);
        ++g1;
      };
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        TLR(1, 1000);
        tc::mbar_wait(&bars[ln ? B_XS_READY : B_XS_FULL], it & 1);
        tc::fence_after_sync();
        TLR(1, 1001);
        for (int c = 0; c < LOOK; ++c) mma1(c);
        for (int c = 0; c < NC; ++c) {
          if (c + LOOK < NC) {
            mma1(c + LOOK);
            if (c + LOOK == NC - 1) tc::umma_commit(&bars[B_XS_EMPTY]);  // every first product of this tile has read the x tile
          }
          const int b = c & 1;
          tc::mbar_wait(&bars[B_HS_FULL0 + b], (g2 >> 1) & 1);
          if (c == 0) tc::mbar_wait(&bars[B_ACC2_EMPTY], (it & 1) ^ 1);
          tc::fence_after_sync();
          TLR(1, 200 + c);
          const uint32_t hs = tc::smem_u32(smem + OFF_HS + b * (BM * 128));
          const uint32_t w2 = tc::smem_u32(smem + OFF_W2 + c * W2C_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc::umma_bf16_ss(tmem_base + COL_ACC2, tc::smem_desc_sw128(hs + k * 32, 16, 1024),
                             tc::smem_desc_sw128(w2 + k * 32, 16, 1024), idesc2, (c | k) != 0 ? 1u : 0u);
          tc::umma_commit(&bars[B_HS_EMPTY0 + b]);
          if (c == NC - 1) tc::umma_commit(&bars[B_ACC2_FULL]);
          ++g2;
        }
      }
    }
  } else {
    // ===================== epilogue teams =====================
    const int team = (warp - 2) >> 3;        // = accumulator / hidden buffer of its chunks
    const int q = warp & 3;                  // TMEM lane quarter
    const int half = ((warp - 2) & 7) >> 2;  // column half of a hidden chunk; column group of the output boxes
    const int r = q * 32 + lane;             // row inside the tile
    const bool leader = ((warp - 2) & 3) == 0 && lane == 0;  // first warp of a column group
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t sw64 = (uint32_t)((r >> 1) & 3);
    const int nmy = half == 0 ? 2 : 1;       // output boxes of this column group: half, half + 2
    // where the 32 columns of box `box` of this row live in the x tile: 16-byte chunk k of the box
    auto xs_chunk = [&](int box, int k) -> unsigned char* {
      return box < 2 ? smem + OFF_XS + r * 128 + ((((uint32_t)(box * 4 + k)) ^ (uint32_t)(r & 7)) << 4)
                     : smem + OFF_XS + BM * 128 + r * 64 + (((uint32_t)k ^ sw64) << 4);
    };
    int it = 0;
    int tl_n = (warp == 2 + 8 * TL_TEAM && lane == 0) ? 0 : 4096;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int m0 = t * BM;
      TLR(2, 1000);
      if (team == 0 && ln) {
        // ---- LayerNorm prologue: two threads per row (32 + 32 | 32 columns), statistics exchanged through shared memory
        tc::mbar_wait_hot(&bars[B_XS_FULL], it & 1);
        uint4 raw[2][4];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 2; ++i)
          if (i < nmy) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              raw[i][k] = *reinterpret_cast<const uint4*>(xs_chunk(half + 2 * i, k));
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw[i][k]);
#pragma unroll
              for (int u = 0; u < 4; ++u) { const float2 f = __bfloat1622float2(h2[u]); s += f.x + f.y; }
            }
          }
        xch[half * BM + r] = s;
        named_bar_sync(3, 32 * TEAM_WARPS);
        const float mean = (s + xch[(half ^ 1) * BM + r]) * (1.f / C);
        float qq = 0.f;
#pragma unroll
        for (int i = 0; i < 2; ++i)
          if (i < nmy) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw[i][k]);
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float2 f = __bfloat1622float2(h2[u]);
                qq = fmaf(f.x - mean, f.x - mean, qq);
                qq = fmaf(f.y - mean, f.y - mean, qq);
              }
            }
          }
        xch[(2 + half) * BM + r] = qq;
        named_bar_sync(3, 32 * TEAM_WARPS);
        const float rstd = rsqrtf((qq + xch[(2 + (half ^ 1)) * BM + r]) * (1.f / C) + p.eps);
#pragma unroll
        for (int i = 0; i < 2; ++i)
          if (i < nmy) {
            const int box = half + 2 * i;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw[i][k]);
              const float4 ga = __ldg(reinterpret_cast<const float4*>(p.gamma + box * 32 + k * 8));
              const float4 gb = __ldg(reinterpret_cast<const float4*>(p.gamma + box * 32 + k * 8 + 4));
              const float4 ba = __ldg(reinterpret_cast<const float4*>(p.beta + box * 32 + k * 8));
              const float4 bb = __ldg(reinterpret_cast<const float4*>(p.beta + box * 32 + k * 8 + 4));
              const float gm[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
              const float bt[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
              uint32_t o[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float2 f = __bfloat1622float2(h2[u]);
                o[u] = pack_bf16((f.x - mean) * rstd * gm[2 * u] + bt[2 * u], (f.y - mean) * rstd * gm[2 * u + 1] + bt[2 * u + 1]);
              }
              *reinterpret_cast<uint4*>(xs_chunk(box, k)) = make_uint4(o[0], o[1], o[2], o[3]);
            }
          }
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars[B_XS_READY]);
        TLR(2, 1001);
      }
      // ---- this team's hidden chunks: acc1 -> + b1 -> GELU -> bf16 A operand of the second product
      for (int c = team; c < NC; c += 2) {
        const uint32_t u = (uint32_t)(it * (NC / 2) + (c >> 1));  // use index of the team's hidden buffer
        const uint32_t gc = (uint32_t)(it * NC + c);              // global chunk index -> first-product accumulator
        const int ab = gc & (NA1 - 1);
        float bv[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.b1 + c * HC + half * 32 + j));
          bv[j] = b4.x; bv[j + 1] = b4.y; bv[j + 2] = b4.z; bv[j + 3] = b4.w;
        }
        TLR(2, 100 + c);
        tc::mbar_wait_hot(&bars[B_ACC1_FULL0 + ab], (gc / NA1) & 1);
        tc::fence_after_sync();
        TLR(2, 200 + c);
        float v[32];
        tc::tmem_ld32(lane_addr + (uint32_t)(ab * HC + half * 32), v);
        tc::tmem_ld_wait();
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars[B_ACC1_EMPTY0 + ab]);
        TLR(2, 300 + c);
        float2* v2 = reinterpret_cast<float2*>(v);
        const float2* b2 = reinterpret_cast<const float2*>(bv);
#pragma unroll
        for (int j = 0; j < 16; ++j) v2[j] = gelu2(add2(v2[j], b2[j]));
        TLR(2, 400 + c);
        tc::mbar_wait_hot(&bars[B_HS_EMPTY0 + team], (u & 1) ^ 1);  // the team's previous second product has read the buffer
        TLR(2, 500 + c);
        unsigned char* hrow = smem + OFF_HS + team * (BM * 128) + r * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint4 o = {pack_bf16(v[k * 8], v[k * 8 + 1]), pack_bf16(v[k * 8 + 2], v[k * 8 + 3]),
                           pack_bf16(v[k * 8 + 4], v[k * 8 + 5]), pack_bf16(v[k * 8 + 6], v[k * 8 + 7])};
          *reinterpret_cast<uint4*>(hrow + ((((uint32_t)(half * 4 + k)) ^ (uint32_t)(r & 7)) << 4)) = o;
        }
        tc::fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars[B_HS_FULL0 + team]);
        TLR(2, 600 + c);
      }
      if (team == 1) {
        // ---- output: acc2 + b2 (+ residual rows, re-read through L2) -> bf16 -> staging slot -> TMA store
        unsigned char* sbase = smem + OFF_SLOT + half * SLOT_BYTES;
        const bool row_ok = m0 + r < p.M;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (i < nmy) {
            const int box = half + 2 * i;
            uint4 rs[4];
            if (p.residual) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                rs[k] = row_ok ? __ldg(reinterpret_cast<const uint4*>(p.residual + (int64_t)(m0 + r) * NOUT + box * BOX_N) + k)
                               : make_uint4(0u, 0u, 0u, 0u);
            }
            if (i == 0) {
              tc::mbar_wait_hot(&bars[B_ACC2_FULL], it & 1);
              tc::fence_after_sync();
              TLR(2, 2000);
            }
            float v[32];
            tc::tmem_ld32(lane_addr + COL_ACC2 + (uint32_t)(box * BOX_N), v);
            tc::tmem_ld_wait();
            if (i == nmy - 1) {  // this warp's rows of the accumulator are in registers: the next tile may overwrite it
              tc::fence_before_sync();
              __syncwarp();
              if (lane == 0) tc::mbar_arrive(&bars[B_ACC2_EMPTY]);
            }
            float2* v2 = reinterpret_cast<float2*>(v);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {  // b2: L1-resident broadcast loads
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.b2 + box * BOX_N + j));
              v2[j / 2] = add2(v2[j / 2], make_float2(b4.x, b4.y));
              v2[j / 2 + 1] = add2(v2[j / 2 + 1], make_float2(b4.z, b4.w));
            }
            if (p.residual) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const __nv_bfloat162* gp = reinterpret_cast<const __nv_bfloat162*>(&rs[k]);
#pragma unroll
                for (int u = 0; u < 4; ++u) v2[k * 4 + u] = add2(v2[k * 4 + u], __bfloat1622float2(gp[u]));
              }
            }
            // the slot's previous store (the group's previous box) has been read: the leader waited, the barrier tells
            if (leader) bulk_wait_read_all();
            named_bar_sync(1 + half, 128);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint4 o = {pack_bf16(v[k * 8], v[k * 8 + 1]), pack_bf16(v[k * 8 + 2], v[k * 8 + 3]),
                               pack_bf16(v[k * 8 + 4], v[k * 8 + 5]), pack_bf16(v[k * 8 + 6], v[k * 8 + 7])};
              *reinterpret_cast<uint4*>(sbase + r * 64 + ((k ^ sw64) << 4)) = o;
            }
            tc::fence_proxy_async();
            named_bar_sync(1 + half, 128);
            if (leader) {
              tma_store_2d(&tmap_out, sbase, box * BOX_N, m0);
              bulk_commit();
            }
          }
        }
        TLR(2, 2001);
      }
    }
    if (team == 1 && leader) bulk_wait_all();
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, r96::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// C = 192, H = 768, NOUT = 192 (second stage: 6 337 tokens per clip).  590 KB of weights do not fit: they stream from L2
// through two rings of two stages (W1 chunk [64, 192] = 24 KB, W2 chunk [192, 64] = 24 KB), requested by TWO producer
// warps (one per ring: the first product runs up to three chunks ahead of the second, so the rings drain at different
// times), and the MMA issuer POLLS both of its streams (next first product / next second product) and issues whichever
// has its operands -- a blocking wait on one stream would stall the other behind a TMA round trip.  Same two epilogue
// teams as above; the twelve chunks alternate between them, the six output boxes are split four / two.
namespace s192 {
constexpr int C = 192, H = 768, NOUT = 192, NC = H / HC, KA = C / 64;     // 12 chunks, 3 K boxes
constexpr int XS_BYTES = KA * BM * 128;                                    // 48 KB
constexpr int W1C_BYTES = KA * HC * 128;                                   // 24 KB
constexpr int W2C_BYTES = NOUT * 128;                                      // 24 KB
constexpr int NW = 2;                                                      // stages of each weight ring
constexpr int OFF_XS = 0;
constexpr int OFF_W1 = OFF_XS + XS_BYTES;
constexpr int OFF_W2 = OFF_W1 + NW * W1C_BYTES;
constexpr int OFF_HS = OFF_W2 + NW * W2C_BYTES;
constexpr int OFF_SLOT = OFF_HS + 2 * BM * 128;                            // 4 x 8 KB: one per (team, column group)
constexpr int OFF_BAR = OFF_SLOT + 4 * SLOT_BYTES;
constexpr int TOTAL = OFF_BAR + 256 + 1024;
constexpr int NA1 = 4;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t COL_ACC2 = NA1 * HC;
constexpr int TEAM_WARPS = 8;
constexpr int NTHR = 64 + 2 * TEAM_WARPS * 32 + 32;                        // 608: + the second producer warp
constexpr int NBOX = NOUT / BOX_N;                                         // 6
static_assert(TOTAL <= SMEM_LIMIT, "shared memory budget");
enum { B_XS_FULL = 0, B_XS_EMPTY, B_W1_FULL0, B_W1_EMPTY0 = B_W1_FULL0 + NW, B_W2_FULL0 = B_W1_EMPTY0 + NW,
       B_W2_EMPTY0 = B_W2_FULL0 + NW, B_ACC1_FULL0 = B_W2_EMPTY0 + NW, B_ACC1_EMPTY0 = B_ACC1_FULL0 + NA1,
       B_HS_FULL0 = B_ACC1_EMPTY0 + NA1, B_HS_FULL1, B_HS_EMPTY0, B_HS_EMPTY1, B_ACC2_FULL, B_ACC2_EMPTY, NUM_BARS };
static_assert(NUM_BARS * 8 + 8 <= 256, "barrier area");
}  // namespace s192

__global__ void __launch_bounds__(s192::NTHR, 1)
mlp_stream192_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w1,
                     const __grid_constant__ CUtensorMap tmap_w2, const __grid_constant__ CUtensorMap tmap_out, ParamsR p) {
  using namespace s192;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + NUM_BARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = (p.M + BM - 1) / BM;
  const int my_tiles = (int)blockIdx.x < num_tiles ? (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const uint32_t total = (uint32_t)my_tiles * NC;  // chunks this CTA processes

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_x); tc::prefetch_tmap(&tmap_w1); tc::prefetch_tmap(&tmap_w2); tc::prefetch_tmap(&tmap_out);
    for (int i = 0; i < NUM_BARS; ++i) {
      int count = 1;
      if ((i >= B_ACC1_EMPTY0 && i < B_ACC1_EMPTY0 + NA1) || i == B_HS_FULL0 || i == B_HS_FULL1) count = TEAM_WARPS;
      if (i == B_ACC2_EMPTY) count = 2 * TEAM_WARPS;
      tc::mbar_init(&bars[i], count);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_ptr, TMEM_COLS);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== producer A: x tiles and the first-product weight ring =====================
    if (tc::elect_one()) {
      uint32_t g = 0;
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int m0 = t * BM;
        tc::mbar_wait(&bars[B_XS_EMPTY], (it & 1) ^ 1);
        tc::mbar_arrive_expect_tx(&bars[B_XS_FULL], XS_BYTES);
#pragma unroll
        for (int kb = 0; kb < KA; ++kb) tc::tma_load_2d(smem + OFF_XS + kb * (BM * 128), &tmap_x, &bars[B_XS_FULL], kb * 64, m0);
        if (t + (int)gridDim.x < num_tiles) {
#pragma unroll
          for (int kb = 0; kb < KA; ++kb) tma_prefetch_l2_2d(&tmap_x, kb * 64, m0 + (int)gridDim.x * BM);
        }
        for (int c = 0; c < NC; ++c, ++g) {
          const int slot = g % NW;
          tc::mbar_wait(&bars[B_W1_EMPTY0 + slot], ((g / NW) & 1) ^ 1);
          unsigned char* w1 = smem + OFF_W1 + slot * W1C_BYTES;
          tc::mbar_arrive_expect_tx(&bars[B_W1_FULL0 + slot], W1C_BYTES);
#pragma unroll
          for (int kb = 0; kb < KA; ++kb) tc::tma_load_2d(w1 + kb * (HC * 128), &tmap_w1, &bars[B_W1_FULL0 + slot], kb * 64, c * HC);
        }
      }
    }
  } else if (warp == 2 + 2 * TEAM_WARPS) {
    // ===================== producer B: the second-product weight ring =====================
    if (tc::elect_one()) {
      for (uint32_t g = 0; g < total; ++g) {
        const int slot = g % NW;
        tc::mbar_wait(&bars[B_W2_EMPTY0 + slot], ((g / NW) & 1) ^ 1);
        tc::mbar_arrive_expect_tx(&bars[B_W2_FULL0 + slot], W2C_BYTES);
        tc::tma_load_2d(smem + OFF_W2 + slot * W2C_BYTES, &tmap_w2, &bars[B_W2_FULL0 + slot], (int)(g % NC) * HC, 0);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: two streams, whichever is ready =====================
    if (tc::elect_one()) {
      constexpr uint32_t idesc1 = tc::idesc_bf16(BM, HC, 0, 0);
      constexpr uint32_t idesc2 = tc::idesc_bf16(BM, NOUT, 0, 0);
      const uint32_t xs = tc::smem_u32(smem + OFF_XS);
      uint32_t i1 = 0, i2 = 0, idle = 0;
      while (i2 < total) {
        bool progressed = false;
        if (i1 < total) {
          const uint32_t c1 = i1 % NC, t1 = i1 / NC;
          const int ws = i1 % NW, ab = i1 % NA1;
          if ((c1 != 0 || tc::mbar_try_wait(&bars[B_XS_FULL], t1 & 1)) &&
              tc::mbar_try_wait(&bars[B_W1_FULL0 + ws], (i1 / NW) & 1) &&
              tc::mbar_try_wait(&bars[B_ACC1_EMPTY0 + ab], ((i1 / NA1) & 1) ^ 1)) {
            tc::fence_after_sync();
            const uint32_t w1 = tc::smem_u32(smem + OFF_W1 + ws * W1C_BYTES);
            const uint32_t d = tmem_base + (uint32_t)(ab * HC);
#pragma unroll
            for (int kb = 0; kb < KA; ++kb)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                tc::umma_bf16_ss(d, tc::smem_desc_sw128(xs + kb * (BM * 128) + k * 32, 16, 1024),
                                 tc::smem_desc_sw128(w1 + kb * (HC * 128) + k * 32, 16, 1024), idesc1, (kb | k) != 0 ? 1u : 0u);
            tc::umma_commit(&bars[B_ACC1_FULL0 + ab]);
            tc::umma_commit(&bars[B_W1_EMPTY0 + ws]);
            if (c1 == NC - 1) tc::umma_commit(&bars[B_XS_EMPTY]);  // every first product of this tile has read the x tile
            ++i1;
            progressed = true;
          }
        }
        if (i2 < i1) {
          const uint32_t c2 = i2 % NC, t2 = i2 / NC;
          const int ws = i2 % NW, b = i2 & 1;
          if (tc::mbar_try_wait(&bars[B_HS_FULL0 + b], (i2 >> 1) & 1) && tc::mbar_try_wait(&bars[B_W2_FULL0 + ws], (i2 / NW) & 1) &&
              (c2 != 0 || tc::mbar_try_wait(&bars[B_ACC2_EMPTY], (t2 & 1) ^ 1))) {
            tc::fence_after_sync();
            const uint32_t hs = tc::smem_u32(smem + OFF_HS + b * (BM * 128));
            const uint32_t w2 = tc::smem_u32(smem + OFF_W2 + ws * W2C_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc::umma_bf16_ss(tmem_base + COL_ACC2, tc::smem_desc_sw128(hs + k * 32, 16, 1024),
                               tc::smem_desc_sw128(w2 + k * 32, 16, 1024), idesc2, (c2 | (uint32_t)k) != 0 ? 1u : 0u);
            tc::umma_commit(&bars[B_W2_EMPTY0 + ws]);
            tc::umma_commit(&bars[B_HS_EMPTY0 + b]);
            if (c2 == NC - 1) tc::umma_commit(&bars[B_ACC2_FULL]);
            ++i2;
            progressed = true;
          }
        }
        if (progressed) {
          idle = 0;
        } else {
          __nanosleep(20);
          if (++idle > (1u << 24)) __trap();  // bounded like every other wait
        }
      }
    }
  } else {
    // ===================== epilogue teams =====================
    const int team = (warp - 2) >> 3;
    const int q = warp & 3;
    const int half = ((warp - 2) & 7) >> 2;
    const int r = q * 32 + lane;
    const bool leader = ((warp - 2) & 3) == 0 && lane == 0;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t sw64 = (uint32_t)((r >> 1) & 3);
    const int grp = team * 2 + half;  // output group: slot, named barrier
    // output boxes of this group: column parity = half, pair index (box >> 1) in {0, 2} for team 0, {1} for team 1
    const int nmy = team == 0 ? 2 : 1;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int m0 = t * BM;
      for (int c = team; c < NC; c += 2) {
        const uint32_t u = (uint32_t)(it * (NC / 2) + (c >> 1));
        const uint32_t gc = (uint32_t)(it * NC + c);
        const int ab = gc % NA1;
        float bv[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.b1 + c * HC + half * 32 + j));
          bv[j] = b4.x; bv[j + 1] = b4.y; bv[j + 2] = b4.z; bv[j + 3] = b4.w;
        }
        tc::mbar_wait_hot(&bars[B_ACC1_FULL0 + ab], (gc / NA1) & 1);
        tc::fence_after_sync();
        float v[32];
        tc::tmem_ld32(lane_addr + (uint32_t)(ab * HC + half * 32), v);
        tc::tmem_ld_wait();
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars[B_ACC1_EMPTY0 + ab]);
        float2* v2 = reinterpret_cast<float2*>(v);
        const float2* b2 = reinterpret_cast<const float2*>(bv);
#pragma unroll
        for (int j = 0; j < 16; ++j) v2[j] = gelu2(add2(v2[j], b2[j]));
        tc::mbar_wait_hot(&bars[B_HS_EMPTY0 + team], (u & 1) ^ 1);
        unsigned char* hrow = smem + OFF_HS + team * (BM * 128) + r * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint4 o = {pack_bf16(v[k * 8], v[k * 8 + 1]), pack_bf16(v[k * 8 + 2], v[k * 8 + 3]),
                           pack_bf16(v[k * 8 + 4], v[k * 8 + 5]), pack_bf16(v[k * 8 + 6], v[k * 8 + 7])};
          *reinterpret_cast<uint4*>(hrow + ((((uint32_t)(half * 4 + k)) ^ (uint32_t)(r & 7)) << 4)) = o;
        }
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars[B_HS_FULL0 + team]);
      }
      // ---- output boxes of this group
      unsigned char* sbase = smem + OFF_SLOT + grp * SLOT_BYTES;
      const bool row_ok = m0 + r < p.M;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (i < nmy) {
          const int box = half + 2 * (team == 0 ? 2 * i : 1);
          uint4 rs[4];
          if (p.residual) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              rs[k] = row_ok ? __ldg(reinterpret_cast<const uint4*>(p.residual + (int64_t)(m0 + r) * NOUT + box * BOX_N) + k)
                             : make_uint4(0u, 0u, 0u, 0u);
          }
          if (i == 0) {
            tc::mbar_wait_hot(&bars[B_ACC2_FULL], it & 1);
            tc::fence_after_sync();
          }
          float v[32];
          tc::tmem_ld32(lane_addr + COL_ACC2 + (uint32_t)(box * BOX_N), v);
          tc::tmem_ld_wait();
          if (i == nmy - 1) {
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bars[B_ACC2_EMPTY]);
          }
          float2* v2 = reinterpret_cast<float2*>(v);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.b2 + box * BOX_N + j));
            v2[j / 2] = add2(v2[j / 2], make_float2(b4.x, b4.y));
            v2[j / 2 + 1] = add2(v2[j / 2 + 1], make_float2(b4.z, b4.w));
          }
          if (p.residual) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const __nv_bfloat162* gp = reinterpret_cast<const __nv_bfloat162*>(&rs[k]);
#pragma unroll
              for (int uu = 0; uu < 4; ++uu) v2[k * 4 + uu] = add2(v2[k * 4 + uu], __bfloat1622float2(gp[uu]));
            }
          }
          if (leader) bulk_wait_read_all();  // the slot's previous store has been read
          named_bar_sync(1 + grp, 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint4 o = {pack_bf16(v[k * 8], v[k * 8 + 1]), pack_bf16(v[k * 8 + 2], v[k * 8 + 3]),
                             pack_bf16(v[k * 8 + 4], v[k * 8 + 5]), pack_bf16(v[k * 8 + 6], v[k * 8 + 7])};
            *reinterpret_cast<uint4*>(sbase + r * 64 + ((k ^ sw64) << 4)) = o;
          }
          tc::fence_proxy_async();
          named_bar_sync(1 + grp, 128);
          if (leader) {
            tma_store_2d(&tmap_out, sbase, box * BOX_N, m0);
            bulk_commit();
          }
        }
      }
    }
    if (leader) bulk_wait_all();
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, s192::TMEM_COLS);
  }
}

// 2-D bf16 tensor [rows, cols], box = [box_rows, 32 cols], 64-byte swizzle
int make_map32(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  svit_tmap_encode_fn enc = svit_get_tmap_encode();
  if (!enc) return SVIT_ENOTSUP;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : SVIT_EINVAL;
}

int launch_resident96(const void* x, const void* w1, const float* b1, const void* w2, const float* b2, const void* residual,
                      void* out, int64_t M, const float* gamma, const float* beta, float eps, cudaStream_t st) {
  using namespace r96;
  CUtensorMap tx64, tx32, tw164, tw132, tw2, tout;
  int rc;
  if ((rc = svit_make_tmap_2d(&tx64, x, (uint64_t)M, C, C, BM))) return rc;
  if ((rc = make_map32(&tx32, x, (uint64_t)M, C, C, BM))) return rc;
  if ((rc = svit_make_tmap_2d(&tw164, w1, (uint64_t)H, C, C, HC))) return rc;
  if ((rc = make_map32(&tw132, w1, (uint64_t)H, C, C, HC))) return rc;
  if ((rc = svit_make_tmap_2d(&tw2, w2, NOUT, (uint64_t)H, (uint64_t)H, NOUT))) return rc;
  if ((rc = make_map32(&tout, out, (uint64_t)M, NOUT, NOUT, BM))) return rc;
  ParamsR p;
  p.b1 = b1; p.b2 = b2; p.gamma = gamma; p.beta = beta; p.eps = eps; p.M = (int)M;
  p.dbg = g_mlp_timeline;
  p.residual = (const bf16*)residual;
  static SvitDevOnce configured;
  if (configured.need()) {
    SVIT_CUDA(cudaFuncSetAttribute(mlp_resident96_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TOTAL));
    configured.done();
  }
  const int64_t tiles = (M + BM - 1) / BM;
  const int grid = (int)(tiles < svit_num_sms() ? tiles : svit_num_sms());
  mlp_resident96_kernel<<<grid, NTHR, TOTAL, st>>>(tx64, tx32, tw164, tw132, tw2, tout, p);
  SVIT_CHECK_LAUNCH();
  return 0;
}

int launch_stream192(const void* x, const void* w1, const float* b1, const void* w2, const float* b2, const void* residual,
                     void* out, int64_t M, cudaStream_t st) {
  using namespace s192;
  CUtensorMap tx, tw1, tw2, tout;
  int rc;
  if ((rc = svit_make_tmap_2d(&tx, x, (uint64_t)M, C, C, BM))) return rc;
  if ((rc = svit_make_tmap_2d(&tw1, w1, (uint64_t)H, C, C, HC))) return rc;
  if ((rc = svit_make_tmap_2d(&tw2, w2, NOUT, (uint64_t)H, (uint64_t)H, NOUT))) return rc;
  if ((rc = make_map32(&tout, out, (uint64_t)M, NOUT, NOUT, BM))) return rc;
  ParamsR p;
  p.dbg = nullptr;
  p.b1 = b1; p.b2 = b2; p.gamma = nullptr; p.beta = nullptr; p.eps = 0.f; p.M = (int)M;
  p.residual = (const bf16*)residual;
  static SvitDevOnce configured;
  if (configured.need()) {
    SVIT_CUDA(cudaFuncSetAttribute(mlp_stream192_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TOTAL));
    configured.done();
  }
  const int64_t tiles = (M + BM - 1) / BM;
  const int grid = (int)(tiles < svit_num_sms() ? tiles : svit_num_sms());
  mlp_stream192_kernel<<<grid, NTHR, TOTAL, st>>>(tx, tw1, tw2, tout, p);
  SVIT_CHECK_LAUNCH();
  return 0;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

// Diagnostic hook (not part of the public header): device buffer of 3 * 4096 * 2 uint64 that CTA 0 of the next launches
// fills with (tag, clock64) events; NULL switches the probe off.
extern "C" int svit_debug_mlp_timeline(void* device_buffer) {
  g_mlp_timeline = (unsigned long long*)device_buffer;
  return 0;
}

extern "C" int svit_mlp_fused_supported(int64_t M, int C, int H, int N) {
  if (M < 1 || M >= (1ll << 31) - BM) return 0;
  if (!((C == 96 && N == 96) || (C == 192 && N == 192))) return 0;
  return H == 4 * C;
}

// out[M, N] = residual + fc2(gelu(fc1(LN(x)))): x [M, C] bf16, w1 [H, C] bf16, w2 [N, H] bf16, b1 [H] / b2 [N] fp32,
// residual [M, N] bf16 or NULL, out [M, N] bf16.  ln_gamma / ln_beta [C] fp32 (both or neither; C = 96 only): x is
// normalised on the fly (LayerNorm over C, eps) before fc1, and `residual` must then be x itself or NULL.
extern "C" int svit_mlp_fused(const void* x, const void* w1, const float* b1, const void* w2, const float* b2,
                              const void* residual, void* out, int64_t M, int C, int H, int N, const float* ln_gamma,
                              const float* ln_beta, float eps, void* stream) {
  if (!x || !w1 || !b1 || !w2 || !b2 || !out || (ln_gamma == nullptr) != (ln_beta == nullptr)) return SVIT_EINVAL;
  if (M == 0) return 0;
  if (!svit_mlp_fused_supported(M, C, H, N)) return SVIT_ENOTSUP;
  if (!aligned16(x) || !aligned16(w1) || !aligned16(w2) || !aligned16(out) || !aligned16(b1) || !aligned16(b2) ||
      (residual && !aligned16(residual)) || (ln_gamma && (!aligned16(ln_gamma) || !aligned16(ln_beta))))
    return SVIT_ENOTSUP;
  if (ln_gamma && (C != 96 || (residual && residual != x))) return SVIT_ENOTSUP;
  if (out == x || out == residual) return SVIT_ENOTSUP;  // rows are re-read (L2 prefetch, residual) while others are stored
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 96) return launch_resident96(x, w1, b1, w2, b2, residual, out, M, ln_gamma, ln_beta, eps, st);
  return launch_stream192(x, w1, b1, w2, b2, residual, out, M, st);
}
