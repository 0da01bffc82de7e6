// Pooled attention on the 5th-generation tensor cores (bf16 operands, fp32 accumulation in TMEM):
//   out = softmax(scale q k^T + decomposed rel-pos bias) v  (+ q on rows >= 1)
// Reference: slowfast/models/attention.py:429-459 with cal_rel_pos_spatial (:84-137) and
// cal_rel_pos_temporal (:140-183) fused into the score tile; the [Nq, Nk] matrix never leaves the SM.
//
// One CTA = 128 query rows of one (batch, head); two CTAs are co-resident per SM so one CTA's softmax
// overlaps the other's MMAs.  192 threads:
//   warp 0    TMA producer (Q tile, rel-pos table passes, K ring of 2, V)
//   warp 1    tcgen05.mma issuer; owns the 256 TMEM columns: S0 | S1 (64 each, double buffered) | O (128)
//   warps 2-5 softmax: thread = query row = TMEM lane
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
constexpr int NTHREADS = 192;
constexpr int STG_PITCH = TP + 1;

constexpr int OFF_Q = 0;                       // 2 boxes x 128 rows x 128 B
constexpr int OFF_K = 32768;                   // 2 stages x (2 boxes x 64 rows x 128 B)
constexpr int OFF_V = OFF_K + 2 * 16384;       // 2 boxes x 64 rows x 128 B
constexpr int OFF_T = OFF_K;                   // tables alias K/V: 2 boxes x 80 rows x 128 B
constexpr int OFF_STG = OFF_K;                 // gather staging aliases K/V: 128 rows x 81 fp32 (41472 B <= 49152)
constexpr int OFF_E = OFF_V + 16384;           // E_s [128][epitch] fp32
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
};

enum {  // barrier slots
  BAR_Q_FULL = 0, BAR_T_FULL, BAR_E_FULL, BAR_E_EMPTY, BAR_K_FULL0, BAR_K_FULL1, BAR_K_EMPTY0, BAR_K_EMPTY1,
  BAR_V_FULL, BAR_V_EMPTY, BAR_S_FULL0, BAR_S_FULL1, BAR_P_FULL0, BAR_P_FULL1, BAR_O_DONE, NUM_BARS
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

template <int KW>
__global__ void __launch_bounds__(NTHREADS, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                   const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_t, Params p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* Es = reinterpret_cast<float*>(smem + OFF_E);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_E + BM * p.epitch * 4);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + NUM_BARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int r0 = blockIdx.x * BM;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_q); tc::prefetch_tmap(&tmap_k); tc::prefetch_tmap(&tmap_v); tc::prefetch_tmap(&tmap_t);
    for (int i = 0; i < NUM_BARS; ++i) {
      const bool four = (i == BAR_E_EMPTY || i == BAR_P_FULL0 || i == BAR_P_FULL1);
      tc::mbar_init(&bars[i], four ? 4 : 1);
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
    if (lane == 0) {
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
      for (int j = 0; j < p.n_tiles; ++j) {
        const int ks = j & 1;
        const int n0 = key_start<KW>(p, j);
        tc::mbar_wait(&bars[BAR_K_EMPTY0 + ks], ((j >> 1) & 1) ^ 1);
        unsigned char* kd = smem + OFF_K + ks * 16384;
        tc::mbar_arrive_expect_tx(&bars[BAR_K_FULL0 + ks], 16384);
        tc::tma_load_3d(kd, &tmap_k, &bars[BAR_K_FULL0 + ks], 0, n0, bh);
        tc::tma_load_3d(kd + 8192, &tmap_k, &bars[BAR_K_FULL0 + ks], 64, n0, bh);
        tc::mbar_wait(&bars[BAR_V_EMPTY], (j & 1) ^ 1);
        tc::mbar_arrive_expect_tx(&bars[BAR_V_FULL], 16384);
        tc::tma_load_3d(smem + OFF_V, &tmap_v, &bars[BAR_V_FULL], 0, n0, bh);
        tc::tma_load_3d(smem + OFF_V + 8192, &tmap_v, &bars[BAR_V_FULL], 64, n0, bh);
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr uint32_t idesc_e = tc::idesc_bf16(BM, TP, 0, 0);
      constexpr uint32_t idesc_s = tc::idesc_bf16(BM, BN, 0, 0);
      constexpr uint32_t idesc_o = tc::idesc_bf16(BM, 128, 0, 1);
      const uint32_t sq = tc::smem_u32(smem + OFF_Q);
      tc::mbar_wait(&bars[BAR_Q_FULL], 0);
      for (int ps = 0; ps < p.n_pass; ++ps) {
        tc::mbar_wait(&bars[BAR_T_FULL], ps & 1);
        tc::fence_after_sync();
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
          tc::mbar_wait(&bars[BAR_V_FULL], i & 1);
          tc::fence_after_sync();
          const uint32_t sv = tc::smem_u32(smem + OFF_V);
#pragma unroll
          for (int k = 0; k < BN / 16; ++k) {
            const uint64_t db = tc::smem_desc_sw128(sv + k * 2048, 8192, 1024);
            tc::umma_bf16_ts(tmem_base + COL_O, tmem_base + COL_S0 + (i & 1) * BN + k * 8, db, idesc_o, (i | k) != 0);
          }
          tc::umma_commit(&bars[BAR_V_EMPTY]);
          tc::umma_commit(&bars[BAR_O_DONE]);
        }
      }
    }
  } else {
    // =========================== softmax warps ===========================
    const int qd = warp & 3;
    const int rl = qd * 32 + lane;          // row within the tile = TMEM lane
    const int row = r0 + rl;                // row within the sequence
    const uint32_t lane_addr = tmem_base + ((uint32_t)(qd * 32) << 16);
    float* Er = Es + rl * p.epitch;
    for (int c = 0; c < p.epitch; ++c) Er[c] = 0.f;
    const bool qpatch = row >= 1 && row <= p.Lq;
    int qi = 0, qj = 0, qt_ = 0;
    if (qpatch) {
      const int pp = row - 1;
      qj = pp % p.qw; qi = (pp / p.qw) % p.qh; qt_ = pp / (p.qw * p.qh);
    }
    // ---- phase E: gather this row's bias terms from the table product
    float* stg = reinterpret_cast<float*>(smem + OFF_STG) + rl * STG_PITCH;
    for (int ps = 0; ps < p.n_pass; ++ps) {
      tc::mbar_wait(&bars[BAR_E_FULL], ps & 1);
      tc::fence_after_sync();
#pragma unroll
      for (int c0 = 0; c0 < TP; c0 += 16) {
        float v[16];
        tc::tmem_ld16(lane_addr + COL_S0 + c0, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) stg[c0 + i] = v[i];
      }
      if (qpatch) {
        const int lo = ps * TP;
        for (int c = 0; c < p.ne; ++c) {
          int g;
          if (c < p.kh) g = p.idx_h[qi * p.kh + c];
          else if (c < p.kh + p.kw) g = p.off_w + p.idx_w[qj * p.kw + (c - p.kh)];
          else g = p.off_t + p.idx_t[qt_ * p.kt + (c - p.kh - p.kw)];
          g -= lo;
          if (g >= 0 && g < TP) Er[c] = stg[g] * 1.4426950408889634f;
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars[BAR_E_EMPTY]);
    }
    // ---- phase S: online softmax over the key tiles
    float ew[KW > 0 ? KW : 1];
    if (KW > 0) {
#pragma unroll
      for (int k = 0; k < (KW > 0 ? KW : 1); ++k) ew[k] = Er[p.kh + k];
    }
    (void)ew;
    float m_ref = -INFINITY, l = 0.f;
    for (int j = 0; j < p.n_tiles; ++j) {
      const int sb = j & 1;
      const int n0 = key_start<KW>(p, j);
      (void)n0;
      tc::mbar_wait(&bars[BAR_S_FULL0 + sb], (j >> 1) & 1);
      tc::fence_after_sync();
      float y[BN];
      tc::tmem_ld32(lane_addr + COL_S0 + sb * BN, y);
      tc::tmem_ld32(lane_addr + COL_S0 + sb * BN + 32, y + 32);
      tc::tmem_ld_wait();
      float mx = -INFINITY;
      if (KW == 0) {
#pragma unroll
        for (int c = 0; c < BN; ++c) {
          const int code = __ldg(p.key_cols + n0 + c);
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
          int gr = j * RPT;
          int tq = gr / p.kh, iq = gr - tq * p.kh;
          float brow[RPT];
#pragma unroll
          for (int r = 0; r < RPT; ++r) {
            brow[r] = (gr + r < p.kthkh) ? Er[p.kh + KWc + tq] + Er[iq] : -INFINITY;
            if (++iq == p.kh) { iq = 0; ++tq; }
          }
          y[0] = j == 0 ? y[0] * p.c1 : -INFINITY;  // cls key (no bias) / duplicated key of the previous tile
          mx = y[0];
#pragma unroll
          for (int c = 1; c < BN; ++c) {
            if (c <= RPT * KWc) {
              y[c] = fmaf(y[c], p.c1, brow[(c - 1) / KWc] + ew[(c - 1) % KWc]);
              mx = fmaxf(mx, y[c]);
            } else {
              y[c] = -INFINITY;
            }
          }
        } else {
          const int nvalid = p.O - (j - p.n_patch_tiles) * BN;
#pragma unroll
          for (int c = 0; c < BN; ++c) {
            y[c] = c < nvalid ? y[c] * p.c1 : -INFINITY;
            mx = fmaxf(mx, y[c]);
          }
        }
      }
      const float m_new = fmaxf(m_ref, mx);
      const bool grow = m_new > m_ref + 8.f;  // lazy rescale: stale reference max is fine while p <= 2^8
      // Observe every O_DONE phase in order (PV_{j-1} has normally finished long before this point): the parity
      // wait only distinguishes "current" from "previous" phase, so no phase may be skipped.
      if (j > 0) {
        tc::mbar_wait(&bars[BAR_O_DONE], (j - 1) & 1);
        tc::fence_after_sync();
      }
      if (__any_sync(0xffffffffu, grow) && j > 0) {
        const float alpha = grow ? exp2f(m_ref - m_new) : 1.f;
#pragma unroll
        for (int c0 = 0; c0 < HD; c0 += 32) {
          float o[32];
          tc::tmem_ld32(lane_addr + COL_O + c0, o);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] *= alpha;
          tc::tmem_st32(lane_addr + COL_O + c0, reinterpret_cast<uint32_t*>(o));
        }
        tc::tmem_st_wait();
        l *= alpha;
      }
      if (grow) m_ref = m_new;
      uint32_t pk[BN / 2];
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < BN; c += 2) {
        const float p0 = exp2f(y[c] - m_ref), p1 = exp2f(y[c + 1] - m_ref);
        sum += p0 + p1;
        pk[c >> 1] = pack2(p0, p1);
      }
      l += sum;
      tc::tmem_st32(lane_addr + COL_S0 + sb * BN, pk);
      tc::tmem_st_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars[BAR_P_FULL0 + sb]);
    }
    // ---- phase O: normalise, residual pooling, store
    tc::mbar_wait(&bars[BAR_O_DONE], (p.n_tiles - 1) & 1);
    tc::fence_after_sync();
    const float inv = 1.f / l;
    const bool valid = row < p.Nq;
    const int b = bh / p.h, head = bh % p.h;
    bf16* op = p.out + (((int64_t)b * p.Nq + row) * p.h + head) * HD;
    const bf16* qp = p.q + ((int64_t)bh * p.Nq + row) * HD;
#pragma unroll
    for (int c0 = 0; c0 < HD; c0 += 32) {
      float o[32];
      tc::tmem_ld32(lane_addr + COL_O + c0, o);
      tc::tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          float r[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) r[u] = o[i + u] * inv;
          if (row >= 1) {
            const uint4 qq = *reinterpret_cast<const uint4*>(qp + c0 + i);
            const __nv_bfloat162* q2 = reinterpret_cast<const __nv_bfloat162*>(&qq);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float2 f = __bfloat1622float2(q2[u]);
              r[2 * u] += f.x;
              r[2 * u + 1] += f.y;
            }
          }
          uint4 w = {pack2(r[0], r[1]), pack2(r[2], r[3]), pack2(r[4], r[5]), pack2(r[6], r[7])};
          *reinterpret_cast<uint4*>(op + c0 + i) = w;
        }
      }
    }
    if (valid && p.lse) p.lse[(int64_t)bh * p.Nq + row] = (m_ref + log2f(l)) * 0.6931471805599453f;
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
  const uint64_t BH = (uint64_t)a->B * a->h;
  CUtensorMap tq, tk, tv, tt;
  int rc;
  if ((rc = svit_make_tmap_3d(&tq, a->q, BH, p.Nq, HD, HD, (uint64_t)p.Nq * HD, BM))) return rc;
  if ((rc = svit_make_tmap_3d(&tk, a->k, BH, p.Nk, HD, HD, (uint64_t)p.Nk * HD, BN))) return rc;
  if ((rc = svit_make_tmap_3d(&tv, a->v, BH, p.Nk, HD, HD, (uint64_t)p.Nk * HD, BN))) return rc;
  if ((rc = svit_make_tmap_2d(&tt, a->rel_tab, p.ntab, HD, HD, TP))) return rc;
  const int smem = OFF_E + BM * p.epitch * 4 + NUM_BARS * 8 + 16 + 1024;
  if (smem > 200 * 1024) return SVIT_ENOTSUP;
  dim3 grid((unsigned)((p.Nq + BM - 1) / BM), (unsigned)BH);
#define ATTN_LAUNCH(KWV)                                                                                              \
  {                                                                                                                   \
    static bool configured = false;                                                                                   \
    if (!configured) {                                                                                                \
      SVIT_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<KWV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
      configured = true;                                                                                              \
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
