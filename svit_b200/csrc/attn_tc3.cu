// Pooled attention on the 5th-generation tensor cores with the decomposed relative-position bias INSIDE the score MMA:
//   out = softmax(scale q k^T + bias) v  (+ q on rows >= 1)           slowfast/models/attention.py:429-459
//   bias[row, key] = q_row . (Rh[i,i'] + Rw[j,j'] + Rt[t,t'])          cal_rel_pos_spatial/temporal (:84-183)
// The bias is a product too: bias = E[row, :] . Sel[:, key], E = the row's gathered table products (<= 31 values)
// and Sel the 0/1 matrix that picks (i', j', t') of a key.  So the score tile is ONE accumulation
//   S = [Q | E'] . [K | Sel]^T        (K dim 96 + 32, E' = E / scale in bf16, column 31 = 1 against a -1e30 mask
//                                      entry of padding keys)
// and the softmax warps see finished logits: per element one FFMA2 half, one MUFU.EX2, a max and a sum.
//
// One CTA = 128 query rows of one (batch, head); two CTAs per SM.  320 threads:
//   warp 0    TMA producer (Q, rel-pos table passes, ring of 2 x {K, Sel}, ring of 2 x V)
//   warp 1    tcgen05.mma issuer; TMEM: S0 | S1 (64 columns each) | O (96)
//   warps 2-9 softmax: thread = query row = TMEM lane; warps w and w+4 split the 64 columns of a score tile
// Phases per CTA:
//   (E)  E_tab = Q . T^T for the concatenated un-gathered tables (passes of 80 rows); each thread gathers 16 of its
//        row's 32 E' columns through the integer index tables and writes them (bf16, 64-byte swizzle) into the
//        E tile next to the Q tile in shared memory.
//   (S)  per 64-key tile: 8 MMAs (4 + 2 + 2 K-steps) -> tcgen05.ld -> online softmax with lazy rescale of O ->
//        P (bf16) written back into the S columns -> O += P V (A from TMEM, V MN-major from smem).
//   (O)  O / l (+ q residual from the resident Q tile) -> bf16 -> coalesced rows of out[b, row, head, :].
// Shared-memory operand layouts: 64-column chunks are 128-byte-swizzled rows, 32-column chunks (Q/K columns 64..95,
// E', Sel) 64-byte-swizzled rows; A and B descriptors carry their own layout type.
#include "tc_common.cuh"
#include "../../include/svit_b200.h"

extern unsigned long long* g_attn_timeline;  // attn_tc.cu (diagnostic hook)

namespace {

constexpr int BM = 128;   // query rows per CTA
constexpr int BN = 64;    // keys per tile
constexpr int HB = 32;    // score columns per softmax thread and tile
constexpr int TP = 80;    // table rows per E pass
constexpr int EK = 32;    // E' / Sel columns of the first chunk (the last one is the mask column)
constexpr int EK2 = 16;   // optional second chunk
constexpr int HD = SVIT_HEAD_DIM;
constexpr int NTHREADS = 320;

constexpr int OFF_Q0 = 0;                 // 128 rows x 128 B (columns 0..63, SW128)
constexpr int OFF_Q1 = 16384;             // 128 rows x 64 B  (columns 64..95, SW64)
constexpr int OFF_ET = 24576;             // 128 rows x 64 B  (E' columns 0..31, SW64)
constexpr int OFF_ET2 = 32768;            // 128 rows x 32 B  (E' columns 32..47, SW32; only when ne > 31)
constexpr int OFF_K = 36864;              // 2 stages x { K0 64 x 128 B | K1 64 x 64 B | Sel 64 x 64 B | Sel2 64 x 32 B }
constexpr int K_STAGE = 18432, K1_OFF = 8192, SEL_OFF = 12288, SEL2_OFF = 16384;
constexpr int V_STAGE = 16384;            // 2 boxes x 64 rows x 128 B
constexpr int OFF_V1 = OFF_K + 2 * K_STAGE;   // V stage 1 sits next to K stage 1: together they are the alias region
constexpr int OFF_V0 = OFF_V1 + V_STAGE;      // V stage 0
constexpr int OFF_ALIAS = OFF_K + K_STAGE;    // K stage 1 + V stage 1 = 34816 B: rel-pos tables, then gather staging
constexpr int OFF_T = OFF_ALIAS;          // tables: 2 boxes x 80 rows x 128 B
constexpr int OFF_STG = OFF_ALIAS;        // gather staging: 128 rows x 49 fp32 (one 48-column half of a table pass)
constexpr int STG_COLS = 48, STG_PITCH = STG_COLS + 1;
constexpr int OFF_X = OFF_V0 + V_STAGE;  // row max / row sum exchange [2][2][128] fp32
constexpr int OFF_BAR = OFF_X + 4 * BM * 4;
constexpr int SMEM_TOTAL = OFF_BAR + 256 + 1024;
constexpr int TMEM_COLS = 256;
constexpr int COL_S0 = 0, COL_O = 128;
static_assert(BM * STG_PITCH * 4 <= K_STAGE + V_STAGE && 2 * TP * 128 <= K_STAGE + V_STAGE, "alias region too small");

struct Params {
  unsigned long long* dbg;
  int h, qh, qw, kh, kw, kt, O;
  int Nq, Nk, Lq, ne;
  int ntab, off_w, off_t, n_pass, n_tiles;
  float c1;         // scale * log2(e)
  float inv_scale;  // 1 / scale
  const int32_t* idx_h;
  const int32_t* idx_w;
  const int32_t* idx_t;
  bf16* out;
  float* lse;
};

// timeline probe (build with SVIT_NVCC_EXTRA=-DSVIT_TIMELINE; tools/attn_timeline.py)
#ifdef SVIT_TIMELINE
#define TL(role, tag)                                                                 \
  do {                                                                                \
    if (p.dbg && blockIdx.x == 1 && blockIdx.y == 0 && tl_n < 4096) {                 \
      p.dbg[((role) * 4096 + tl_n) * 2] = (unsigned long long)(tag);                  \
      p.dbg[((role) * 4096 + tl_n) * 2 + 1] = (unsigned long long)clock64();          \
      ++tl_n;                                                                         \
    }                                                                                 \
  } while (0)
#else
#define TL(role, tag) do { (void)tl_n; } while (0)
#endif

enum {  // barrier slots
  BAR_Q_FULL = 0, BAR_T_FULL, BAR_E_FULL, BAR_E_EMPTY, BAR_E_READY, BAR_K_FULL0, BAR_K_FULL1, BAR_K_EMPTY0, BAR_K_EMPTY1,
  BAR_V_FULL0, BAR_V_FULL1, BAR_V_EMPTY0, BAR_V_EMPTY1, BAR_S_FULL0, BAR_S_FULL1, BAR_P_FULL0, BAR_P_FULL1, BAR_O_DONE,
  BAR_O_FINAL,  // completes exactly once, when the last P.V MMA has retired
  NUM_BARS
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// K-major shared-memory descriptor, 64-byte swizzle: rows of 64 B, 8-row groups 512 B apart
__device__ __forceinline__ uint64_t smem_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(16 >> 4) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}

// K-major, 32-byte swizzle: rows of 32 B, 8-row groups 256 B apart
__device__ __forceinline__ uint64_t smem_desc_sw32(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(16 >> 4) << 16;
  d |= (uint64_t)(256 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;
  return d;
}

// X16: the bias vector has more than 31 entries; entries 31.. live in a second 16-column chunk (one more K-step)
template <bool X16>
__global__ void __launch_bounds__(NTHREADS, 2)
attn_fwd_tc3_kernel(const __grid_constant__ CUtensorMap tmap_q0, const __grid_constant__ CUtensorMap tmap_q1,
                    const __grid_constant__ CUtensorMap tmap_k0, const __grid_constant__ CUtensorMap tmap_k1,
                    const __grid_constant__ CUtensorMap tmap_sel, const __grid_constant__ CUtensorMap tmap_sel2,
                    const __grid_constant__ CUtensorMap tmap_v,
                    const __grid_constant__ CUtensorMap tmap_t, Params p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + NUM_BARS);
  float* xch = reinterpret_cast<float*>(smem + OFF_X);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int r0 = blockIdx.x * BM;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_q0); tc::prefetch_tmap(&tmap_q1); tc::prefetch_tmap(&tmap_k0); tc::prefetch_tmap(&tmap_k1);
    tc::prefetch_tmap(&tmap_sel); tc::prefetch_tmap(&tmap_v); tc::prefetch_tmap(&tmap_t);
    if (X16) tc::prefetch_tmap(&tmap_sel2);
    for (int i = 0; i < NUM_BARS; ++i) {
      const bool eight = (i == BAR_E_EMPTY || i == BAR_E_READY || i == BAR_P_FULL0 || i == BAR_P_FULL1);
      tc::mbar_init(&bars[i], eight ? 8 : 1);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_ptr, TMEM_COLS);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (tc::elect_one()) {
      int tl_n = 0;
      TL(0, 9000);
      tc::mbar_arrive_expect_tx(&bars[BAR_Q_FULL], 16384 + 8192);
      tc::tma_load_3d(smem + OFF_Q0, &tmap_q0, &bars[BAR_Q_FULL], 0, r0, bh);
      tc::tma_load_3d(smem + OFF_Q1, &tmap_q1, &bars[BAR_Q_FULL], 64, r0, bh);
      // K (+ Sel) and V tiles are requested separately: the K tile of key tile j+1 goes out BEFORE the V tile of tile j.
      // A V stage frees up only when the P.V MMA two tiles back has completed (late in a tile period); queueing the
      // next K request behind that wait made every score MMA start ~800 cycles after its operands could have been
      // there (TMA latency ~1.5k cycles; tools/attn_timeline.py).
      auto load_k = [&](int j) {
        const int ks = j & 1;
        const int n0 = j * BN;
        tc::mbar_wait(&bars[BAR_K_EMPTY0 + ks], ((j >> 1) & 1) ^ 1);
        TL(0, 100 + j);
        unsigned char* kd = smem + OFF_K + ks * K_STAGE;
        tc::mbar_arrive_expect_tx(&bars[BAR_K_FULL0 + ks], X16 ? K_STAGE : SEL2_OFF);
        tc::tma_load_3d(kd, &tmap_k0, &bars[BAR_K_FULL0 + ks], 0, n0, bh);
        tc::tma_load_3d(kd + K1_OFF, &tmap_k1, &bars[BAR_K_FULL0 + ks], 64, n0, bh);
        tc::tma_load_2d(kd + SEL_OFF, &tmap_sel, &bars[BAR_K_FULL0 + ks], 0, n0);
        if (X16) tc::tma_load_2d(kd + SEL2_OFF, &tmap_sel2, &bars[BAR_K_FULL0 + ks], EK, n0);
      };
      auto load_v = [&](int j) {
        const int ks = j & 1;
        const int n0 = j * BN;
        tc::mbar_wait(&bars[BAR_V_EMPTY0 + ks], ((j >> 1) & 1) ^ 1);
        unsigned char* vd = smem + (ks ? OFF_V1 : OFF_V0);
        tc::mbar_arrive_expect_tx(&bars[BAR_V_FULL0 + ks], V_STAGE);
        tc::tma_load_3d(vd, &tmap_v, &bars[BAR_V_FULL0 + ks], 0, n0, bh);
        tc::tma_load_3d(vd + 8192, &tmap_v, &bars[BAR_V_FULL0 + ks], 64, n0, bh);
      };
      auto load_kv = [&](int j) { load_k(j); load_v(j); };
      tc::mbar_arrive_expect_tx(&bars[BAR_T_FULL], 2 * TP * 128);
      tc::tma_load_2d(smem + OFF_T, &tmap_t, &bars[BAR_T_FULL], 0, 0);
      tc::tma_load_2d(smem + OFF_T + TP * 128, &tmap_t, &bars[BAR_T_FULL], 64, 0);
      load_kv(0);  // stage 0 is not part of the alias region: the first key tile arrives during phase E
      for (int ps = 1; ps < p.n_pass; ++ps) {
        tc::mbar_wait(&bars[BAR_E_EMPTY], (ps - 1) & 1);
        tc::mbar_arrive_expect_tx(&bars[BAR_T_FULL], 2 * TP * 128);
        tc::tma_load_2d(smem + OFF_T, &tmap_t, &bars[BAR_T_FULL], 0, ps * TP);
        tc::tma_load_2d(smem + OFF_T + TP * 128, &tmap_t, &bars[BAR_T_FULL], 64, ps * TP);
      }
      tc::mbar_wait(&bars[BAR_E_EMPTY], (p.n_pass - 1) & 1);  // tables + staging alias stage 1 of the K/V ring
      if (p.n_tiles > 1) load_k(1);
      for (int j = 1; j < p.n_tiles; ++j) {
        if (j + 1 < p.n_tiles) load_k(j + 1);
        load_v(j);
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (tc::elect_one()) {
      constexpr uint32_t idesc_e = tc::idesc_bf16(BM, TP, 0, 0);
      constexpr uint32_t idesc_s = tc::idesc_bf16(BM, BN, 0, 0);
      constexpr uint32_t idesc_o = tc::idesc_bf16(BM, HD, 0, 1);
      const uint32_t sq0 = tc::smem_u32(smem + OFF_Q0), sq1 = tc::smem_u32(smem + OFF_Q1), se = tc::smem_u32(smem + OFF_ET);
      const uint32_t se2 = tc::smem_u32(smem + OFF_ET2);
      (void)se2;
      int tl_n = 0;
      tc::mbar_wait(&bars[BAR_Q_FULL], 0);
      TL(1, 9001);
      for (int ps = 0; ps < p.n_pass; ++ps) {
        tc::mbar_wait(&bars[BAR_T_FULL], ps & 1);
        tc::fence_after_sync();
        const uint32_t st = tc::smem_u32(smem + OFF_T);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc::umma_bf16_ss(tmem_base + COL_S0, tc::smem_desc_sw128(sq0 + k * 32, 16, 1024),
                           tc::smem_desc_sw128(st + k * 32, 16, 1024), idesc_e, k != 0);
#pragma unroll
        for (int k = 0; k < 2; ++k)
          tc::umma_bf16_ss(tmem_base + COL_S0, smem_desc_sw64(sq1 + k * 32),
                           tc::smem_desc_sw128(st + TP * 128 + k * 32, 16, 1024), idesc_e, 1u);
        tc::umma_commit(&bars[BAR_E_FULL]);
      }
      tc::mbar_wait(&bars[BAR_E_READY], 0);  // E' tile written; E_tab columns are about to become S0/S1
      tc::fence_after_sync();
      TL(1, 9002);
      for (int j = 0; j <= p.n_tiles; ++j) {
        if (j < p.n_tiles) {
          const int ks = j & 1;
#ifdef SVIT_TIMELINE
          TL(1, 800 + j);
          TL(1, tc::mbar_try_wait(&bars[BAR_K_FULL0 + ks], (j >> 1) & 1) ? 9901 : 9900);
#endif
          tc::mbar_wait(&bars[BAR_K_FULL0 + ks], (j >> 1) & 1);
          tc::fence_after_sync();
          TL(1, 100 + j);
          const uint32_t sk = tc::smem_u32(smem + OFF_K + ks * K_STAGE);
          const uint32_t d = tmem_base + COL_S0 + (j & 1) * BN;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc::umma_bf16_ss(d, tc::smem_desc_sw128(sq0 + k * 32, 16, 1024), tc::smem_desc_sw128(sk + k * 32, 16, 1024),
                             idesc_s, k != 0);
#pragma unroll
          for (int k = 0; k < 2; ++k)
            tc::umma_bf16_ss(d, smem_desc_sw64(sq1 + k * 32), smem_desc_sw64(sk + K1_OFF + k * 32), idesc_s, 1u);
#pragma unroll
          for (int k = 0; k < 2; ++k)
            tc::umma_bf16_ss(d, smem_desc_sw64(se + k * 32), smem_desc_sw64(sk + SEL_OFF + k * 32), idesc_s, 1u);
          if (X16) tc::umma_bf16_ss(d, smem_desc_sw32(se2), smem_desc_sw32(sk + SEL2_OFF), idesc_s, 1u);
          TL(1, 600 + j);
          tc::umma_commit(&bars[BAR_K_EMPTY0 + ks]);
          tc::umma_commit(&bars[BAR_S_FULL0 + (j & 1)]);
          TL(1, 700 + j);
        }
        if (j >= 1) {
          const int i = j - 1;
          tc::mbar_wait(&bars[BAR_P_FULL0 + (i & 1)], (i >> 1) & 1);
          TL(1, 200 + i);
          tc::mbar_wait(&bars[BAR_V_FULL0 + (i & 1)], (i >> 1) & 1);
          tc::fence_after_sync();
          TL(1, 300 + i);
          const uint32_t sv = tc::smem_u32(smem + ((i & 1) ? OFF_V1 : OFF_V0));
#pragma unroll
          for (int k = 0; k < BN / 16; ++k) {
            const uint64_t db = tc::smem_desc_sw128(sv + k * 2048, 8192, 1024);
            tc::umma_bf16_ts(tmem_base + COL_O, tmem_base + COL_S0 + (i & 1) * BN + k * 8, db, idesc_o, (i | k) != 0);
          }
          TL(1, 400 + i);
          tc::umma_commit(&bars[BAR_V_EMPTY0 + (i & 1)]);
          tc::umma_commit(&bars[BAR_O_DONE]);
          TL(1, 500 + i);
          if (i == p.n_tiles - 1) tc::umma_commit(&bars[BAR_O_FINAL]);
        }
      }
    }
  } else {
    // =========================== softmax warps ===========================
    const int qd = warp & 3;
    const int half = (warp - 2) >> 2;
    const int rl = qd * 32 + lane;          // row within the tile = TMEM lane
    const int row = r0 + rl;                // row within the sequence
    const uint32_t lane_addr = tmem_base + ((uint32_t)(qd * 32) << 16);
    const int pair_bar = 1 + qd;
    const bool qpatch = row >= 1 && row <= p.Lq;
    int tl_n = (warp == 2 && lane == 0) ? 0 : 4096;
    TL(2, 9000);
    // ---- phase E: this thread owns E' columns [16 half, 16 half + 16) of the first chunk (column 31 = mask) and, with
    // X16, columns [8 half, 8 half + 8) of the second; bias entry e sits in column e (e < 31) or 32 + (e - 31)
    constexpr int NG = X16 ? 24 : 16;
    int g[NG];
    {
      int qi = 0, qj = 0, qt_ = 0;
      if (qpatch) {
        const int pp = row - 1;
        qj = pp % p.qw; qi = (pp / p.qw) % p.qh; qt_ = pp / (p.qw * p.qh);
      }
#pragma unroll
      for (int u = 0; u < NG; ++u) {
        const int c = u < 16 ? half * 16 + u : (EK - 1) + half * 8 + (u - 16);  // bias entry
        g[u] = -1;
        if (qpatch && !(u < 16 && c == EK - 1)) {
          if (c < p.kh) g[u] = __ldg(p.idx_h + qi * p.kh + c);
          else if (c < p.kh + p.kw) g[u] = p.off_w + __ldg(p.idx_w + qj * p.kw + (c - p.kh));
          else if (c < p.ne) g[u] = p.off_t + __ldg(p.idx_t + qt_ * p.kt + (c - p.kh - p.kw));
        }
      }
    }
    float ev[NG];
#pragma unroll
    for (int u = 0; u < NG; ++u) ev[u] = 0.f;
    float* stg = reinterpret_cast<float*>(smem + OFF_STG) + rl * STG_PITCH;
    for (int ps = 0; ps < p.n_pass; ++ps) {
      tc::mbar_wait_hot(&bars[BAR_E_FULL], ps & 1);
      tc::fence_after_sync();
      const int lo = ps * TP;
#pragma unroll
      for (int hb = 0; hb < TP; hb += STG_COLS) {  // the 80 columns of a pass are staged 48 + 32 at a time
        if (hb > 0) named_bar_sync(pair_bar, 64);  // both warps are done with the previous half
#pragma unroll
        for (int c0 = hb; c0 < TP && c0 < hb + STG_COLS; c0 += 16) {
          if (((c0 >> 4) & 1) == half) {  // warp-uniform: alternate 16-column chunks
            float v[16];
            tc::tmem_ld16(lane_addr + COL_S0 + c0, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) stg[c0 - hb + i] = v[i];
          }
        }
        named_bar_sync(pair_bar, 64);  // the staged half row is complete
#pragma unroll
        for (int u = 0; u < NG; ++u) {
          const int gg = g[u] - lo - hb;
          if (g[u] >= 0 && gg >= 0 && gg < STG_COLS && gg + hb < TP) ev[u] = stg[gg];
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars[BAR_E_EMPTY]);
    }
    {
#pragma unroll
      for (int u = 0; u < NG; ++u) ev[u] *= p.inv_scale;
      if (half == 1) ev[15] = 1.0f;  // column 31: multiplies the mask row of Sel (every query row, patch or not)
      const uint32_t sw = (uint32_t)((rl >> 1) & 3);
      unsigned char* erow = smem + OFF_ET + rl * 64;
      const uint4 lo4 = {pack2(ev[0], ev[1]), pack2(ev[2], ev[3]), pack2(ev[4], ev[5]), pack2(ev[6], ev[7])};
      const uint4 hi4 = {pack2(ev[8], ev[9]), pack2(ev[10], ev[11]), pack2(ev[12], ev[13]), pack2(ev[14], ev[15])};
      *reinterpret_cast<uint4*>(erow + (((uint32_t)(half * 2) ^ sw) << 4)) = lo4;
      *reinterpret_cast<uint4*>(erow + (((uint32_t)(half * 2 + 1) ^ sw) << 4)) = hi4;
      if (X16) {
        const uint4 x4 = {pack2(ev[NG - 8], ev[NG - 7]), pack2(ev[NG - 6], ev[NG - 5]), pack2(ev[NG - 4], ev[NG - 3]),
                          pack2(ev[NG - 2], ev[NG - 1])};
        *reinterpret_cast<uint4*>(smem + OFF_ET2 + rl * 32 + (((uint32_t)half ^ (uint32_t)((rl >> 2) & 1)) << 4)) = x4;
      }
      tc::fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars[BAR_E_READY]);
    }
    // ---- phase S: online softmax over the key tiles; the accumulator already holds (s + bias / scale)
    float m_ref = -INFINITY, l = 0.f;
    const float2 c1c1 = make_float2(p.c1, p.c1);
    TL(2, 9001);
    for (int j = 0; j < p.n_tiles; ++j) {
      const int sb = j & 1;
      tc::mbar_wait_hot(&bars[BAR_S_FULL0 + sb], (j >> 1) & 1);
      tc::fence_after_sync();
      TL(2, 100 + j);
      float y[HB];
      tc::tmem_ld32(lane_addr + COL_S0 + sb * BN + half * HB, y);
      tc::tmem_ld_wait();
      TL(2, 200 + j);
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // four independent max chains
#pragma unroll
      for (int c = 0; c < HB; c += 4) {
        m4[0] = fmaxf(m4[0], y[c]); m4[1] = fmaxf(m4[1], y[c + 1]);
        m4[2] = fmaxf(m4[2], y[c + 2]); m4[3] = fmaxf(m4[3], y[c + 3]);
      }
      float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      // row max over both halves (double-buffered exchange slot); the barrier also orders this tile's score reads
      // of BOTH warps before either overwrites score columns with P
      xch[((j & 1) * 2 + half) * BM + rl] = mx;
      named_bar_sync(pair_bar, 64);
      mx = fmaxf(mx, xch[((j & 1) * 2 + (half ^ 1)) * BM + rl]) * p.c1;  // c1 > 0: max commutes with the scaling
      TL(2, 300 + j);
      const float m_new = fmaxf(m_ref, mx);
      const bool grow = m_new > m_ref + 8.f;  // lazy rescale: a stale reference max is fine while p <= 2^8
      // O is only touched when some row of the warp needs a rescale.  Having seen S_FULL(j) implies PV_{j-2} is done
      // (MMAs retire in issue order), so the parity wait for phase j-1 is unambiguous even if earlier phases were
      // never observed.
      if (__any_sync(0xffffffffu, grow) && j > 0) {  // both warps of the quarter take the same decision
        tc::mbar_wait_hot(&bars[BAR_O_DONE], (j - 1) & 1);
        tc::fence_after_sync();
        const float alpha = grow ? tc::ex2_approx(m_ref - m_new) : 1.f;
        const uint32_t oaddr = lane_addr + COL_O + half * (HD / 2);
        float o[32];
        tc::tmem_ld32(oaddr, o);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] *= alpha;
        tc::tmem_st32(oaddr, reinterpret_cast<uint32_t*>(o));
        tc::tmem_ld16(oaddr + 32, o);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] *= alpha;
        tc::tmem_st16(oaddr + 32, reinterpret_cast<uint32_t*>(o));
        tc::tmem_st_wait();
        l *= alpha;
      }
      if (grow) m_ref = m_new;
      TL(2, 400 + j);
      uint32_t pk[HB / 2];
      float2 sum2 = make_float2(0.f, 0.f);
      {
        const float2 negm = make_float2(-m_ref, -m_ref);
        const float2* yy = reinterpret_cast<const float2*>(y);
        // columns of this half that hold real keys: only the last tile is ragged (Nk = 457 leaves 9 of 64).  Padding
        // keys carry the -1e30 mask entry, so their probability is exactly 0: write the zero without spending the
        // special-function unit on it (the loop is MUFU-bound; 55 of 512 exponentials per row at Nk = 457).
        const int nv = p.Nk - j * BN - half * HB;
        if (nv >= HB) {
#pragma unroll
          for (int c = 0; c < HB / 2; ++c) {
            const float2 d = tc::fma2(yy[c], c1c1, negm);
            float2 e;
            e.x = tc::ex2_approx(d.x);
            e.y = tc::ex2_approx(d.y);
            sum2 = tc::add2(sum2, e);
            pk[c] = pack2(e.x, e.y);
          }
        } else {
#pragma unroll
          for (int c = 0; c < HB / 2; ++c) {
            pk[c] = 0u;
            if (2 * c < nv) {  // warp-uniform
              const float2 d = tc::fma2(yy[c], c1c1, negm);
              float2 e;
              e.x = tc::ex2_approx(d.x);
              e.y = 2 * c + 1 < nv ? tc::ex2_approx(d.y) : 0.f;
              sum2 = tc::add2(sum2, e);
              pk[c] = pack2(e.x, e.y);
            }
          }
        }
      }
      l += sum2.x + sum2.y;
      TL(2, 500 + j);
      tc::tmem_st16(lane_addr + COL_S0 + sb * BN + half * (HB / 2), pk);
      tc::tmem_st_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars[BAR_P_FULL0 + sb]);
      TL(2, 600 + j);
    }
    // ---- phase O: normalise, residual pooling, store
    // The last P.V product is observed through its own single-use barrier.  O_DONE cannot serve here: at this point
    // between n_tiles-2 and n_tiles of its phases may be complete, and a parity wait only tells the current phase from
    // the previous one -- a warp that got here after the last P.V had already retired would wait on parity
    // (n_tiles-2)&1 = n_tiles&1, i.e. for a phase that never completes.
    tc::mbar_wait_hot(&bars[BAR_O_FINAL], 0);
    tc::fence_after_sync();
    TL(2, 9003);
    xch[((p.n_tiles & 1) * 2 + half) * BM + rl] = l;  // partial row sums of the two halves
    named_bar_sync(pair_bar, 64);
    l += xch[((p.n_tiles & 1) * 2 + (half ^ 1)) * BM + rl];
    const float inv = 1.f / l;
    const int b = bh / p.h, head = bh % p.h;
    // O rows are staged in the (now idle) K/V ring, 208-byte pitch: conflict-free 16-byte row writes; the two warps
    // of a quarter then stream its 32 rows out as contiguous 192-byte segments (12 lanes per row)
    constexpr int OPITCH = 208;
    unsigned char* ostg = smem + OFF_K + (qd * 32) * OPITCH;
    {
      const int cbase = half * (HD / 2);  // 48 columns per warp: 32 + 16
      float o[48];
      tc::tmem_ld32(lane_addr + COL_O + cbase, o);
      tc::tmem_ld16(lane_addr + COL_O + cbase + 32, o + 32);
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 48; i += 8) {
        float r[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) r[u] = o[i + u] * inv;
        if (row >= 1) {
          // residual pooling: the Q tile is still resident in shared memory
          const int col = cbase + i;
          const unsigned char* qsrc = col < 64
              ? smem + OFF_Q0 + rl * 128 + ((((uint32_t)col >> 3) ^ (uint32_t)(rl & 7)) << 4)
              : smem + OFF_Q1 + rl * 64 + ((((uint32_t)(col - 64) >> 3) ^ (uint32_t)((rl >> 1) & 3)) << 4);
          const uint4 qq = *reinterpret_cast<const uint4*>(qsrc);
          const __nv_bfloat162* q2 = reinterpret_cast<const __nv_bfloat162*>(&qq);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float2 f = __bfloat1622float2(q2[u]);
            r[2 * u] += f.x;
            r[2 * u + 1] += f.y;
          }
        }
        const uint4 w = {pack2(r[0], r[1]), pack2(r[2], r[3]), pack2(r[4], r[5]), pack2(r[6], r[7])};
        *reinterpret_cast<uint4*>(ostg + lane * OPITCH + (cbase + i) * 2) = w;
      }
    }
    named_bar_sync(pair_bar, 64);
    {
      const int rbase = r0 + qd * 32;
      bf16* obase = p.out + (((int64_t)b * p.Nq + rbase) * p.h + head) * HD;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int idx = i * 64 + half * 32 + lane;
        const int rr = idx / 12, ch = idx - rr * 12;
        if (rbase + rr < p.Nq) {
          const uint4 w = *reinterpret_cast<const uint4*>(ostg + rr * OPITCH + ch * 16);
          *reinterpret_cast<uint4*>(obase + (int64_t)rr * p.h * HD + ch * 8) = w;
        }
      }
    }
    TL(2, 9004);
    if (half == 0 && row < p.Nq && p.lse) p.lse[(int64_t)bh * p.Nq + row] = (m_ref + log2f(l)) * 0.6931471805599453f;
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// bf16 tensor map with a 32-column box and 64-byte swizzle (second K chunk of Q / K, the Sel table)
int make_map32(CUtensorMap* m, const void* ptr, int rank, uint64_t d2, uint64_t rows, uint64_t cols, uint64_t ld_row,
               uint64_t ld_d2, uint32_t box_rows) {
  svit_tmap_encode_fn enc = svit_get_tmap_encode();
  if (!enc) return SVIT_ENOTSUP;
  cuuint64_t dims[3] = {cols, rows, d2};
  cuuint64_t strides[2] = {ld_row * 2, ld_d2 * 2};
  cuuint32_t box[3] = {32, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : SVIT_EINVAL;
}

// 16-column box, 32-byte swizzle (second chunk of the Sel table)
int make_map16(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  svit_tmap_encode_fn enc = svit_get_tmap_encode();
  if (!enc) return SVIT_ENOTSUP;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {16, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : SVIT_EINVAL;
}

}  // namespace

// Requires the Sel table (a->sel_tab): 32 columns for kh + kw + kt <= 31, 48 columns for up to 47 bias entries.
int svit_attn_tc3_supported(const svit_attn_args* a) {
  if (a->dtype != SVIT_BF16) return 0;
  if (!a->rel_tab || !a->idx_h || !a->idx_w || !a->idx_t || !a->sel_tab) return 0;
  const int ne = a->kh + a->kw + a->kt;
  if (!((a->sel_cols == EK && ne <= EK - 1) || (a->sel_cols == EK + EK2 && ne <= EK - 1 + EK2))) return 0;
  if (!aligned16(a->q) || !aligned16(a->k) || !aligned16(a->v) || !aligned16(a->out) || !aligned16(a->rel_tab) ||
      !aligned16(a->sel_tab))
    return 0;
  return 1;
}

int svit_attn_fwd_tc3(const svit_attn_args* a, cudaStream_t st) {
  Params p;
  p.h = a->h; p.qh = a->qh; p.qw = a->qw; p.kh = a->kh; p.kw = a->kw; p.kt = a->kt; p.O = a->O;
  p.Lq = a->qt * a->qh * a->qw;
  p.Nq = 1 + p.Lq + a->O;
  p.Nk = 1 + a->kt * a->kh * a->kw + a->O;
  p.ne = a->kh + a->kw + a->kt;
  p.ntab = a->ntab_h + a->ntab_w + a->ntab_t;
  p.off_w = a->ntab_h;
  p.off_t = a->ntab_h + a->ntab_w;
  p.n_pass = (p.ntab + TP - 1) / TP;
  p.n_tiles = (p.Nk + BN - 1) / BN;
  p.c1 = a->scale * 1.4426950408889634f;
  p.inv_scale = 1.0f / a->scale;
  p.idx_h = a->idx_h; p.idx_w = a->idx_w; p.idx_t = a->idx_t;
  p.out = (bf16*)a->out; p.lse = a->lse;
  p.dbg = g_attn_timeline;
  const uint64_t BH = (uint64_t)a->B * a->h;
  CUtensorMap tq0, tq1, tk0, tk1, tsel, tsel2, tv, tt;
  const bool x16 = a->sel_cols == EK + EK2;
  int rc;
  if ((rc = svit_make_tmap_3d(&tq0, a->q, BH, p.Nq, HD, HD, (uint64_t)p.Nq * HD, BM))) return rc;
  if ((rc = make_map32(&tq1, a->q, 3, BH, p.Nq, HD, HD, (uint64_t)p.Nq * HD, BM))) return rc;
  if ((rc = svit_make_tmap_3d(&tk0, a->k, BH, p.Nk, HD, HD, (uint64_t)p.Nk * HD, BN))) return rc;
  if ((rc = make_map32(&tk1, a->k, 3, BH, p.Nk, HD, HD, (uint64_t)p.Nk * HD, BN))) return rc;
  if ((rc = make_map32(&tsel, a->sel_tab, 2, 1, (uint64_t)p.n_tiles * BN, (uint64_t)a->sel_cols, (uint64_t)a->sel_cols, 0, BN))) return rc;
  tsel2 = tsel;
  if (x16 && (rc = make_map16(&tsel2, a->sel_tab, (uint64_t)p.n_tiles * BN, (uint64_t)a->sel_cols, BN))) return rc;
  if ((rc = svit_make_tmap_3d(&tv, a->v, BH, p.Nk, HD, HD, (uint64_t)p.Nk * HD, BN))) return rc;
  if ((rc = svit_make_tmap_2d(&tt, a->rel_tab, p.ntab, HD, HD, TP))) return rc;
  dim3 grid((unsigned)((p.Nq + BM - 1) / BM), (unsigned)BH);
  static SvitDevOnce configured;
  if (configured.need()) {
    SVIT_CUDA(cudaFuncSetAttribute(attn_fwd_tc3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    SVIT_CUDA(cudaFuncSetAttribute(attn_fwd_tc3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    configured.done();
  }
  if (x16) attn_fwd_tc3_kernel<true><<<grid, NTHREADS, SMEM_TOTAL, st>>>(tq0, tq1, tk0, tk1, tsel, tsel2, tv, tt, p);
  else attn_fwd_tc3_kernel<false><<<grid, NTHREADS, SMEM_TOTAL, st>>>(tq0, tq1, tk0, tk1, tsel, tsel2, tv, tt, p);
  SVIT_CHECK_LAUNCH();
  return 0;
}
