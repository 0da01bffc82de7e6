// Pooled attention on the 5th-generation tensor cores (bf16 operands, fp32 accumulation in TMEM):
//   out = softmax(scale q k^T + decomposed rel-pos bias) v  (+ q on rows >= 1)
// Reference: slowfast/models/attention.py:429-459 with cal_rel_pos_spatial (:84-137) and
// cal_rel_pos_temporal (:140-183) fused into the score tile; the [Nq, Nk] matrix never leaves the SM.
//
// One CTA = 128 query rows of one (batch, head); two CTAs are co-resident per SM so one CTA's softmax
// overlaps the other's MMAs.  320 threads:
//   warp 0    TMA producer (Q tile, rel-pos table passes, K ring of 2, V ring of 2)
//   warp 1    tcgen05.mma issuer; owns the 256 TMEM columns: S0 | S1 (64 each, double buffered) | O (128)
//   warps 2-9 softmax: thread = query row = TMEM lane; warps w and w+4 split the 64 columns of a score tile
// Phases per CTA:
//   (E)  E_tab = Q . T^T  for the concatenated un-gathered tables T (passes of 80 rows, one MMA each);
//        every thread picks the kh + kw + kt entries its (t, i, j) selects through the integer index
//        tables -> E_s[row][.] in shared memory (pre-multiplied by log2 e).
//   (S)  per 64-key tile: S = Q K^T (SS MMA) -> tcgen05.ld -> y = s*scale*log2e + E_h[i'] + E_w[j'] + E_t[t']
//        -> online softmax with lazy rescale of O (only when the row max grows by > 8) -> P (bf16) written
//        back into the S columns (tcgen05.st) -> O += P V (TS MMA: A from TMEM, V tile MN-major from smem).
//   (O)  O / l (+ q residual) -> bf16 -> out[b, row, head, :].
#include "tc_common.cuh"
#include "../../include/svit_b200.h"

namespace {

constexpr int BM = 128;   // query rows per CTA
constexpr int BN = 64;    // keys per tile
constexpr int TP = 80;    // table rows per E pass
constexpr int HD = SVIT_HEAD_DIM;
constexpr int NTHREADS = 320;   // TMA warp, MMA warp, 8 softmax warps
constexpr int HB = BN / 2;      // score columns per softmax thread and tile
constexpr int STG_PITCH = TP + 1;

constexpr int OFF_Q = 0;                       // 2 boxes x 128 rows x 128 B
constexpr int OFF_K = 32768;                   // 2 stages x (2 boxes x 64 rows x 128 B)
constexpr int OFF_V = OFF_K + 2 * 16384;       // 2 stages x (2 boxes x 64 rows x 128 B)
constexpr int OFF_T = OFF_K;                   // tables alias K/V: 2 boxes x 80 rows x 128 B
constexpr int OFF_STG = OFF_K;                 // gather staging aliases K/V: 128 rows x 81 fp32 (41472 B <= 65536); never live with the tables
constexpr int OFF_E = OFF_V + 2 * 16384;       // E_s [128][epitch] fp32
constexpr int TMEM_COLS = 256;
constexpr int COL_S0 = 0, COL_O = 128;

struct Params {
  int h, qh, qw, kh, kw, kt, O;
  int Nq, Nk, Lq, ne, epitch;
  int ntab, off_w, off_t, n_pass, n_tiles;
  int n_patch_tiles, kthkh, Lk;  // fast path: key tiles aligned to rows of the key grid
  float c1;  // scale * log2(e)
  const int32_t* idx_h;
  const int32_t* idx_w;
  const int32_t* idx_t;
  const int32_t* key_cols;
  const bf16* q;
  bf16* out;
  float* lse;
  unsigned long long* dbg;  // optional timeline buffer (CTA (0,0) only)
};

// timeline probe (build with SVIT_NVCC_EXTRA=-DSVIT_TIMELINE; tools/attn_timeline.py)
#ifdef SVIT_TIMELINE
#define TL(role, tag)                                                                 \
  do {                                                                                \
    if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && tl_n < 4096) {                 \
      p.dbg[((role) * 4096 + tl_n) * 2] = (unsigned long long)(tag);                  \
      p.dbg[((role) * 4096 + tl_n) * 2 + 1] = (unsigned long long)clock64();          \
      ++tl_n;                                                                         \
    }                                                                                 \
  } while (0)
#else
#define TL(role, tag) do { (void)tl_n; } while (0)
#endif

enum {  // barrier slots
  BAR_Q_FULL = 0, BAR_T_FULL, BAR_E_FULL, BAR_E_EMPTY, BAR_K_FULL0, BAR_K_FULL1, BAR_K_EMPTY0, BAR_K_EMPTY1,
  BAR_V_FULL0, BAR_V_FULL1, BAR_V_EMPTY0, BAR_V_EMPTY1, BAR_S_FULL0, BAR_S_FULL1, BAR_P_FULL0, BAR_P_FULL1, BAR_O_DONE, NUM_BARS
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// First key row of tile j.  Generic path (KW == 0): 64 consecutive keys.  Fast path (KW == kw known at compile
// time): patch tile j = [one extra leading key | RPT rows of the key grid]; the leading key is cls for j == 0 and
// masked otherwise; the object keys follow in tiles of 64.  Column c of a patch tile is then (row (c-1)/KW,
// col (c-1)%KW) with COMPILE-TIME row/col, so the bias costs one FADD per score instead of index decoding.
template <int KW>
__device__ __forceinline__ int key_start(const Params& p, int j) {
  if (KW == 0) return j * BN;
  constexpr int RPT = KW > 0 ? 63 / (KW > 0 ? KW : 1) : 1;
  return j < p.n_patch_tiles ? j * RPT * KW : 1 + p.Lk + (j - p.n_patch_tiles) * BN;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// y = y * c1 + bias for the HALF-th 32 columns of a patch-key tile.  Tile column c >= 1 is key (row (c-1)/KW,
// col (c-1)%KW) of the tile, so with KW and HALF known at compile time every bias is brow[const] + ew[const];
// column 0 carries b0 (cls key of tile 0 / masked duplicate), columns past the last whole key row are masked.
template <int KW, int HALF>
__device__ __forceinline__ void patch_bias(float* y, const float* brow, const float* ew, float c1, float b0) {
  constexpr int RPT = 63 / KW;
  float2* y2 = reinterpret_cast<float2*>(y);
  const float2 c1c1 = make_float2(c1, c1);
#pragma unroll
  for (int i = 0; i < HB; i += 2) {
    const int c = HALF * HB + i;
    float2 bb;
    bb.x = c == 0 ? b0 : (c <= RPT * KW ? brow[(c - 1) / KW] + ew[(c - 1) % KW] : -INFINITY);
    bb.y = (c + 1 <= RPT * KW) ? brow[c / KW] + ew[c % KW] : -INFINITY;
    y2[i >> 1] = tc::fma2(y2[i >> 1], c1c1, bb);
  }
}

template <int KW>
__global__ void __launch_bounds__(NTHREADS, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                   const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_t, Params p) {
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment by pointer offset (not an integer round trip) so accesses stay in the shared state space
  unsigned char* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  float* Es = reinterpret_cast<float*>(smem + OFF_E);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_E + BM * p.epitch * 4);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + NUM_BARS);
  float* xch = reinterpret_cast<float*>(bars + NUM_BARS + 2);  // [2 slots][2 halves][BM] row max / row sum exchange

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int r0 = blockIdx.x * BM;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_q); tc::prefetch_tmap(&tmap_k); tc::prefetch_tmap(&tmap_v); tc::prefetch_tmap(&tmap_t);
    for (int i = 0; i < NUM_BARS; ++i) {
      const bool four = (i == BAR_E_EMPTY || i == BAR_P_FULL0 || i == BAR_P_FULL1);
      tc::mbar_init(&bars[i], four ? 8 : 1);
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
      tc::mbar_arrive_expect_tx(&bars[BAR_Q_FULL], 32768);
      tc::tma_load_3d(smem + OFF_Q, &tmap_q, &bars[BAR_Q_FULL], 0, r0, bh);
      tc::tma_load_3d(smem + OFF_Q + 16384, &tmap_q, &bars[BAR_Q_FULL], 64, r0, bh);
      for (int ps = 0; ps < p.n_pass; ++ps) {
        if (ps > 0) tc::mbar_wait(&bars[BAR_E_EMPTY], (ps - 1) & 1);
        tc::mbar_arrive_expect_tx(&bars[BAR_T_FULL], 2 * TP * 128);
        tc::tma_load_2d(smem + OFF_T, &tmap_t, &bars[BAR_T_FULL], 0, ps * TP);
        tc::tma_load_2d(smem + OFF_T + TP * 128, &tmap_t, &bars[BAR_T_FULL], 64, ps * TP);
      }
      tc::mbar_wait(&bars[BAR_E_EMPTY], (p.n_pass - 1) & 1);  // tables + staging alias the K/V buffers
      // the K request of tile j+1 goes out before the V request of tile j (a V stage frees up late in a tile period;
      // see attn_tc3.cu)
      auto load_k = [&](int j) {
        const int ks = j & 1;
        const int n0 = key_start<KW>(p, j);
        tc::mbar_wait(&bars[BAR_K_EMPTY0 + ks], ((j >> 1) & 1) ^ 1);
        TL(0, 100 + j);
        unsigned char* kd = smem + OFF_K + ks * 16384;
        tc::mbar_arrive_expect_tx(&bars[BAR_K_FULL0 + ks], 16384);
        tc::tma_load_3d(kd, &tmap_k, &bars[BAR_K_FULL0 + ks], 0, n0, bh);
        tc::tma_load_3d(kd + 8192, &tmap_k, &bars[BAR_K_FULL0 + ks], 64, n0, bh);
      };
      auto load_v = [&](int j) {
        const int ks = j & 1;
        const int n0 = key_start<KW>(p, j);
        tc::mbar_wait(&bars[BAR_V_EMPTY0 + ks], ((j >> 1) & 1) ^ 1);
        TL(0, 200 + j);
        unsigned char* vd = smem + OFF_V + ks * 16384;
        tc::mbar_arrive_expect_tx(&bars[BAR_V_FULL0 + ks], 16384);
        tc::tma_load_3d(vd, &tmap_v, &bars[BAR_V_FULL0 + ks], 0, n0, bh);
        tc::tma_load_3d(vd + 8192, &tmap_v, &bars[BAR_V_FULL0 + ks], 64, n0, bh);
      };
      load_k(0);
      for (int j = 0; j < p.n_tiles; ++j) {
        if (j + 1 < p.n_tiles) load_k(j + 1);
        load_v(j);
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (tc::elect_one()) {
      constexpr uint32_t idesc_e = tc::idesc_bf16(BM, TP, 0, 0);
      constexpr uint32_t idesc_s = tc::idesc_bf16(BM, BN, 0, 0);
      constexpr uint32_t idesc_o = tc::idesc_bf16(BM, HD, 0, 1);  // N = 96: the zero-padded half atom of V is not multiplied
      const uint32_t sq = tc::smem_u32(smem + OFF_Q);
      int tl_n = 0;
      tc::mbar_wait(&bars[BAR_Q_FULL], 0);
      TL(1, 9001);
      for (int ps = 0; ps < p.n_pass; ++ps) {
        tc::mbar_wait(&bars[BAR_T_FULL], ps & 1);
        tc::fence_after_sync();
        TL(1, 9500 + ps);
        const uint32_t st = tc::smem_u32(smem + OFF_T);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          const uint64_t da = tc::smem_desc_sw128(sq + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
          const uint64_t db = tc::smem_desc_sw128(st + (k >> 2) * (TP * 128) + (k & 3) * 32, 16, 1024);
          tc::umma_bf16_ss(tmem_base + COL_S0, da, db, idesc_e, k != 0);
        }
        tc::umma_commit(&bars[BAR_E_FULL]);
      }
      tc::mbar_wait(&bars[BAR_E_EMPTY], (p.n_pass - 1) & 1);  // E_tab columns are about to become S0/S1
      tc::fence_after_sync();
      for (int j = 0; j <= p.n_tiles; ++j) {
        if (j < p.n_tiles) {
          const int ks = j & 1;
          tc::mbar_wait(&bars[BAR_K_FULL0 + ks], (j >> 1) & 1);
          tc::fence_after_sync();
          TL(1, 100 + j);
          const uint32_t sk = tc::smem_u32(smem + OFF_K + ks * 16384);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) {
            const uint64_t da = tc::smem_desc_sw128(sq + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
            const uint64_t db = tc::smem_desc_sw128(sk + (k >> 2) * 8192 + (k & 3) * 32, 16, 1024);
            tc::umma_bf16_ss(tmem_base + COL_S0 + (j & 1) * BN, da, db, idesc_s, k != 0);
          }
          tc::umma_commit(&bars[BAR_K_EMPTY0 + ks]);
          tc::umma_commit(&bars[BAR_S_FULL0 + (j & 1)]);
        }
        if (j >= 1) {
          const int i = j - 1;
          tc::mbar_wait(&bars[BAR_P_FULL0 + (i & 1)], (i >> 1) & 1);
          TL(1, 200 + i);
          tc::mbar_wait(&bars[BAR_V_FULL0 + (i & 1)], (i >> 1) & 1);
          tc::fence_after_sync();
          TL(1, 300 + i);
          const uint32_t sv = tc::smem_u32(smem + OFF_V + (i & 1) * 16384);
#pragma unroll
          for (int k = 0; k < BN / 16; ++k) {
            const uint64_t db = tc::smem_desc_sw128(sv + k * 2048, 8192, 1024);
            tc::umma_bf16_ts(tmem_base + COL_O, tmem_base + COL_S0 + (i & 1) * BN + k * 8, db, idesc_o, (i | k) != 0);
          }
          tc::umma_commit(&bars[BAR_V_EMPTY0 + (i & 1)]);
          tc::umma_commit(&bars[BAR_O_DONE]);
        }
      }
    }
  } else {
    // =========================== softmax warps ===========================
    // Eight warps.  Warps w and w + 4 share TMEM lane quarter qd = w & 3 (32 query rows, thread = row); `half`
    // selects the 32-column half of every 64-key score tile, the parity class of the E columns in phase E and the
    // 48-column half of O in the epilogue.  The two warps of a quarter meet on named barrier 1 + qd (64 threads).
    const int qd = warp & 3;
    const int half = (warp - 2) >> 2;
    const int rl = qd * 32 + lane;          // row within the tile = TMEM lane
    const int row = r0 + rl;                // row within the sequence
    const uint32_t lane_addr = tmem_base + ((uint32_t)(qd * 32) << 16);
    const int pair_bar = 1 + qd;
    float* Er = Es + rl * p.epitch;
    for (int c = half; c < p.epitch; c += 2) Er[c] = 0.f;
    const bool qpatch = row >= 1 && row <= p.Lq;
    int qi = 0, qj = 0, qt_ = 0;
    if (qpatch) {
      const int pp = row - 1;
      qj = pp % p.qw; qi = (pp / p.qw) % p.qh; qt_ = pp / (p.qw * p.qh);
    }
    int tl_n = (warp == 2 && lane == 0) ? 0 : 4096;
    // ---- phase E: gather this row's bias terms from the table product
    float* stg = reinterpret_cast<float*>(smem + OFF_STG) + rl * STG_PITCH;
    for (int ps = 0; ps < p.n_pass; ++ps) {
      TL(2, 9050 + ps);
      tc::mbar_wait_hot(&bars[BAR_E_FULL], ps & 1);
      tc::fence_after_sync();
      TL(2, 9100 + ps);
#pragma unroll
      for (int c0 = 0; c0 < TP; c0 += 16) {
        if (((c0 >> 4) & 1) == half) {  // warp-uniform: alternate 16-column chunks
          float v[16];
          tc::tmem_ld16(lane_addr + COL_S0 + c0, v);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) stg[c0 + i] = v[i];
        }
      }
      named_bar_sync(pair_bar, 64);  // the staged row is complete
      TL(2, 9200 + ps);
      if (qpatch) {
        const int lo = ps * TP;
        // E columns of this half's parity; index loads are issued eight at a time before the staging reads
        for (int c0 = half; c0 < p.ne; c0 += 16) {
          int g[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int c = c0 + 2 * u;
            g[u] = -1;
            if (c < p.kh) g[u] = __ldg(p.idx_h + qi * p.kh + c);
            else if (c < p.kh + p.kw) g[u] = p.off_w + __ldg(p.idx_w + qj * p.kw + (c - p.kh));
            else if (c < p.ne) g[u] = p.off_t + __ldg(p.idx_t + qt_ * p.kt + (c - p.kh - p.kw));
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int gg = g[u] - lo;
            if (g[u] >= 0 && gg >= 0 && gg < TP) Er[c0 + 2 * u] = stg[gg] * 1.4426950408889634f;
          }
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars[BAR_E_EMPTY]);
    }
    named_bar_sync(pair_bar, 64);  // both parity classes of this row's E vector are written
    // ---- phase S: online softmax over the key tiles
    float ew[KW > 0 ? KW : 1];
    if (KW > 0) {
#pragma unroll
      for (int k = 0; k < (KW > 0 ? KW : 1); ++k) ew[k] = Er[p.kh + k];
    }
    (void)ew;
    float m_ref = -INFINITY, l = 0.f;
    int tq0 = 0, iq0 = 0;
    (void)tq0; (void)iq0;
    TL(2, 9002);
    for (int j = 0; j < p.n_tiles; ++j) {
      const int sb = j & 1;
      const int n0 = key_start<KW>(p, j);
      (void)n0;
      tc::mbar_wait_hot(&bars[BAR_S_FULL0 + sb], (j >> 1) & 1);
      tc::fence_after_sync();
      TL(2, 100 + j);
      float y[HB];
      tc::tmem_ld32(lane_addr + COL_S0 + sb * BN + half * HB, y);
      tc::tmem_ld_wait();
      TL(2, 200 + j);
      float mx = -INFINITY;
      if (KW == 0) {
#pragma unroll
        for (int c = 0; c < HB; ++c) {
          const int code = __ldg(p.key_cols + n0 + half * HB + c);
          const float bias = Er[code & 0xff] + Er[(code >> 8) & 0xff] + Er[(code >> 16) & 0xff];
          float v = fmaf(y[c], p.c1, bias);
          v = code < 0 ? -INFINITY : v;
          y[c] = v;
          mx = fmaxf(mx, v);
        }
      } else {
        constexpr int KWc = KW > 0 ? KW : 1;
        constexpr int RPT = 63 / KWc;
        if (j < p.n_patch_tiles) {
          const int gr = j * RPT;
          int tq = tq0, iq = iq0;  // (t', i') of key-grid row gr, carried across tiles (no division)
          float brow[RPT];
#pragma unroll
          for (int r = 0; r < RPT; ++r) {
            brow[r] = (gr + r < p.kthkh) ? Er[p.kh + KWc + tq] + Er[iq] : -INFINITY;
            if (++iq == p.kh) { iq = 0; ++tq; }
          }
          tq0 = tq; iq0 = iq;
          const float b0 = j == 0 ? 0.f : -INFINITY;
          if (half == 0) patch_bias<KWc, 0>(y, brow, ew, p.c1, b0);
          else patch_bias<KWc, 1>(y, brow, ew, p.c1, b0);
        } else {
          const int nvalid = p.O - (j - p.n_patch_tiles) * BN - half * HB;
          float2* y2 = reinterpret_cast<float2*>(y);
          const float2 c1c1 = make_float2(p.c1, p.c1);
#pragma unroll
          for (int c = 0; c < HB; c += 2) {
            const float2 bb = make_float2(c < nvalid ? 0.f : -INFINITY, c + 1 < nvalid ? 0.f : -INFINITY);
            y2[c >> 1] = tc::fma2(y2[c >> 1], c1c1, bb);
          }
        }
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // four independent max chains
#pragma unroll
        for (int c = 0; c < HB; c += 4) {
          m4[0] = fmaxf(m4[0], y[c]); m4[1] = fmaxf(m4[1], y[c + 1]);
          m4[2] = fmaxf(m4[2], y[c + 2]); m4[3] = fmaxf(m4[3], y[c + 3]);
        }
        mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      }
      // row max over both halves (double-buffered exchange slot); the barrier also orders this tile's score reads
      // of BOTH warps before either overwrites score columns with P
      xch[((j & 1) * 2 + half) * BM + rl] = mx;
      named_bar_sync(pair_bar, 64);
      mx = fmaxf(mx, xch[((j & 1) * 2 + (half ^ 1)) * BM + rl]);
      TL(2, 300 + j);
      const float m_new = fmaxf(m_ref, mx);
      const bool grow = m_new > m_ref + 8.f;  // lazy rescale: stale reference max is fine while p <= 2^8
      // Observe every O_DONE phase in order (PV_{j-1} has normally finished long before this point): the parity
      // wait only distinguishes "current" from "previous" phase, so no phase may be skipped.
      if (j > 0) {
        tc::mbar_wait_hot(&bars[BAR_O_DONE], (j - 1) & 1);
        tc::fence_after_sync();
      }
      if (__any_sync(0xffffffffu, grow) && j > 0) {  // both warps of the quarter take the same decision
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
#pragma unroll
        for (int c = 0; c < HB / 2; ++c) {
          const float2 d = tc::add2(yy[c], negm);
          float2 e;
          e.x = tc::ex2_approx(d.x);
          e.y = tc::ex2_approx(d.y);
          sum2 = tc::add2(sum2, e);
          pk[c] = pack2(e.x, e.y);
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
    tc::mbar_wait_hot(&bars[BAR_O_DONE], (p.n_tiles - 1) & 1);
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
          // residual pooling: the Q tile is still resident in shared memory (128-byte swizzled boxes of 64 columns)
          const int col = cbase + i;
          const uint4 qq = *reinterpret_cast<const uint4*>(smem + OFF_Q + (col >> 6) * 16384 + rl * 128 +
                                                           ((((col & 63) >> 3) ^ (rl & 7)) << 4));
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

}  // namespace

unsigned long long* g_attn_timeline = nullptr;  // shared with attn_tc3.cu

// Diagnostic hook (not part of the public header): see tools/attn_timeline.py
extern "C" int svit_debug_attn_timeline(void* device_buffer) {
  g_attn_timeline = (unsigned long long*)device_buffer;
  return 0;
}

int svit_attn_tc_supported(const svit_attn_args* a) {
  if (a->dtype != SVIT_BF16) return 0;
  if (!a->rel_tab || !a->idx_h || !a->idx_w || !a->idx_t || !a->key_cols) return 0;
  const int ne = a->kh + a->kw + a->kt;
  if (ne > 200) return 0;
  if (!aligned16(a->q) || !aligned16(a->k) || !aligned16(a->v) || !aligned16(a->out) || !aligned16(a->rel_tab)) return 0;
  return 1;
}

int svit_attn_fwd_tc(const svit_attn_args* a, cudaStream_t st) {
  Params p;
  p.h = a->h; p.qh = a->qh; p.qw = a->qw; p.kh = a->kh; p.kw = a->kw; p.kt = a->kt; p.O = a->O;
  p.Lq = a->qt * a->qh * a->qw;
  p.Nq = 1 + p.Lq + a->O;
  p.Nk = 1 + a->kt * a->kh * a->kw + a->O;
  p.ne = a->kh + a->kw + a->kt;
  p.epitch = (p.ne + 1) | 1;  // odd pitch: conflict-free column reads; slot `ne` stays 0 (cls / object keys)
  p.ntab = a->ntab_h + a->ntab_w + a->ntab_t;
  p.off_w = a->ntab_h;
  p.off_t = a->ntab_h + a->ntab_w;
  p.n_pass = (p.ntab + TP - 1) / TP;
  p.n_tiles = (p.Nk + BN - 1) / BN;
  p.Lk = a->kt * a->kh * a->kw;
  p.kthkh = a->kt * a->kh;
  p.n_patch_tiles = 0;
  const int KWs = (a->kw == 7 || a->kw == 14 || a->kw == 10 || a->kw == 20) ? a->kw : 0;
  if (KWs) {
    const int rpt = 63 / KWs;
    p.n_patch_tiles = (p.kthkh + rpt - 1) / rpt;
    p.n_tiles = p.n_patch_tiles + (a->O + BN - 1) / BN;
  }
  p.c1 = a->scale * 1.4426950408889634f;
  p.idx_h = a->idx_h; p.idx_w = a->idx_w; p.idx_t = a->idx_t; p.key_cols = a->key_cols;
  p.q = (const bf16*)a->q; p.out = (bf16*)a->out; p.lse = a->lse;
  p.dbg = g_attn_timeline;
  const uint64_t BH = (uint64_t)a->B * a->h;
  CUtensorMap tq, tk, tv, tt;
  int rc;
  if ((rc = svit_make_tmap_3d(&tq, a->q, BH, p.Nq, HD, HD, (uint64_t)p.Nq * HD, BM))) return rc;
  if ((rc = svit_make_tmap_3d(&tk, a->k, BH, p.Nk, HD, HD, (uint64_t)p.Nk * HD, BN))) return rc;
  if ((rc = svit_make_tmap_3d(&tv, a->v, BH, p.Nk, HD, HD, (uint64_t)p.Nk * HD, BN))) return rc;
  if ((rc = svit_make_tmap_2d(&tt, a->rel_tab, p.ntab, HD, HD, TP))) return rc;
  const int smem = OFF_E + BM * p.epitch * 4 + NUM_BARS * 8 + 16 + 4 * BM * 4 + 1024;
  if (smem > 200 * 1024) return SVIT_ENOTSUP;
  dim3 grid((unsigned)((p.Nq + BM - 1) / BM), (unsigned)BH);
#define ATTN_LAUNCH(KWV)                                                                                              \
  {                                                                                                                   \
    static SvitDevOnce configured;                                                                                   \
    if (configured.need()) {                                                                                                \
      SVIT_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<KWV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
      configured.done();                                                                                              \
    }                                                                                                                 \
    attn_fwd_tc_kernel<KWV><<<grid, NTHREADS, smem, st>>>(tq, tk, tv, tt, p);                                         \
  }
  switch (KWs) {
    case 7: ATTN_LAUNCH(7) break;
    case 14: ATTN_LAUNCH(14) break;
    case 10: ATTN_LAUNCH(10) break;
    case 20: ATTN_LAUNCH(20) break;
    default: ATTN_LAUNCH(0) break;
  }
#undef ATTN_LAUNCH
  SVIT_CHECK_LAUNCH();
  return 0;
}
