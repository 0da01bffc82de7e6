// Attention backward, first half, fused on the tensor cores:  for every 128 x 128 (query x key) tile
//     S  = [q | E'] [k | Sel]^T   and   dP = dO v^T     two tcgen05 accumulations side by side in TMEM
//     P  = exp(scale S - lse)                  the rel-pos bias E[row, i'] + E[row, kh + j'] + E[row, kh + kw + t'] is part
//                                              of the score product, as in the forward kernel (attn_tc3.cu): E' = E / scale
//                                              as a bf16 hi + lo pair (written by the prep kernel: fp32-accurate, the
//                                              backward differentiates the exact bias), Sel the 0/1 key-selection matrix.  The
//                                              epilogue used to gather the three terms per element from shared memory
//                                              (ncu: 31 instructions per element, half of the kernel's issue slots).
//     dS = P (dP - delta)
// and only the bf16 operands of the second round of GEMMs (P for dV, dS for dK / dQ / dE) go to HBM -- the fp32 score
// and dP matrices never exist in memory (attention.py:429-459 differentiated; see attn_bwd_tc.cu for the whole plan).
//
// Persistent, warp-specialised, one CTA per SM (320 threads), same skeleton as gemm_tc.cu:
//   warp 0    TMA producer: 4-stage ring of {A 128 x 64, B 128 x 64} bf16 tiles (128-byte swizzle); per tile six
//             stage loads: (q, k) columns 0..63 and 64..95, (E'hi, Sel), (E'lo, Sel), then (dO, v) columns 0..63 and 64..95.  dO is read in place from the
//             head-merged [B, Nq, h, 96] gradient through a 4-D tensor map (sample, head) coordinate.
//   warp 1    MMA issuer: 6 + 2 nep/16 + 6 tcgen05.mma 128 x 128 x 16 per tile; accumulators S | dP (2 x 128 columns), double
//             buffered (512 TMEM columns) so the MMAs of tile i+1 overlap the epilogue of tile i
//   warps 2-9 epilogue, thread = query row: tcgen05.ld 32 columns of S and of dP, MUFU.EX2, pack to bf16, P and dS
//             rows leave through a per-warp shared-memory transpose.
#include <cstdlib>

#include "tc_common.cuh"
#include "../../include/svit_b200.h"

namespace {

constexpr int BM = 128, BN = 128, BK = 64, STAGES = 4;
constexpr int HD = SVIT_HEAD_DIM;
constexpr int NUM_THREADS = 320, EPI_WARPS = 8;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr uint32_t TMEM_COLS = 512;
constexpr int MAX_KEYS_PADDED = 4096;  // key -> column code table (4 B per key) lives in shared memory
constexpr int OPITCH = 80;             // staging pitch of a 32-column bf16 row segment

struct Params {
  int B, h, Nq, Nk, Nkp, Lq, Lk, kh, kw, kt, nep;
  int m_tiles, n_tiles;
  int stage_out;  // 1: P / dS leave through a per-warp shared-memory transpose (coalesced row segments)
  float sc;  // scale * log2(e)
  const float* lse;
  const float* ws_delta;
  bf16* P;
  bf16* dS;
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
attn_bwd_sdp_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                    const __grid_constant__ CUtensorMap tmap_do, const __grid_constant__ CUtensorMap tmap_v,
                    const __grid_constant__ CUtensorMap tmap_e, const __grid_constant__ CUtensorMap tmap_sel, Params p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  // per epilogue warp two 32-row x 64-byte tiles (P, dS) at an 80-byte pitch: conflict-free 16-byte row writes
  unsigned char* ostg = smem + STAGES * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ostg + (p.stage_out ? EPI_WARPS * 2 * 32 * OPITCH : 0) + 8);
  full_bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(full_bar) + 7) & ~uintptr_t(7));
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_bh = p.m_tiles * p.n_tiles;
  const int num_tiles = per_bh * p.B * p.h;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_q);
    tc::prefetch_tmap(&tmap_k);
    tc::prefetch_tmap(&tmap_do);
    tc::prefetch_tmap(&tmap_v);
    tc::prefetch_tmap(&tmap_e);
    tc::prefetch_tmap(&tmap_sel);
    for (int i = 0; i < STAGES; ++i) {
      tc::mbar_init(&full_bar[i], 1);
      tc::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&tmem_full[i], 1);
      tc::mbar_init(&tmem_empty[i], EPI_WARPS);
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
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < num_tiles; w += gridDim.x) {
        const int bh = w / per_bh, rem = w - bh * per_bh;
        const int m0 = (rem / p.n_tiles) * BM, n0 = (rem % p.n_tiles) * BN;
        const int b = bh / p.h, head = bh - b * p.h;
#pragma unroll
        for (int kb = 0; kb < 6; ++kb) {
          tc::mbar_wait(&empty_bar[stage], phase ^ 1);
          unsigned char* sa = smem + stage * STAGE_BYTES;
          unsigned char* sb = sa + A_BYTES;
          tc::mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          if (kb < 2) {
            tc::tma_load_4d(sa, &tmap_q, &full_bar[stage], kb * BK, m0, 0, bh);
            tc::tma_load_4d(sb, &tmap_k, &full_bar[stage], kb * BK, n0, 0, bh);
          } else if (kb < 4) {  // bias operands: E' hi / lo rows of this (b, head), Sel rows of the keys (shared by all)
            tc::tma_load_4d(sa, &tmap_e, &full_bar[stage], (kb - 2) * p.nep, m0, 0, bh);
            tc::tma_load_4d(sb, &tmap_sel, &full_bar[stage], 0, n0, 0, 0);
          } else {
            tc::tma_load_4d(sa, &tmap_do, &full_bar[stage], (kb - 4) * BK, m0, head, b);
            tc::tma_load_4d(sb, &tmap_v, &full_bar[stage], (kb - 4) * BK, n0, 0, bh);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (tc::elect_one()) {
      constexpr uint32_t idesc = tc::idesc_bf16(BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = blockIdx.x; w < num_tiles; w += gridDim.x, ++it) {
        const int as = it & 1;
        tc::mbar_wait_hot(&tmem_empty[as], ((it >> 1) & 1) ^ 1);
        tc::fence_after_sync();
        const int esteps = p.nep >> 4;  // bias entries, 16 per MMA (nep is a multiple of 16 on this path)
#pragma unroll
        for (int kb = 0; kb < 6; ++kb) {
          tc::mbar_wait_hot(&full_bar[stage], phase);
          tc::fence_after_sync();
          const uint32_t d_tmem = tmem_base + (uint32_t)(as * 2 * BN + (kb >= 4 ? BN : 0));
          const uint32_t sa = tc::smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
          // columns 64..95 only in the second block of q / k and of dO / v
          const int ksteps = (kb == 2 || kb == 3) ? esteps : ((kb == 1 || kb == 5) ? (HD - BK) / 16 : BK / 16);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            if (k < ksteps)
              tc::umma_bf16_ss(d_tmem, tc::smem_desc_sw128(sa + k * 32, 16, 1024), tc::smem_desc_sw128(sb + k * 32, 16, 1024),
                               idesc, ((kb != 0 && kb != 4) || k != 0) ? 1u : 0u);
          tc::umma_commit(&empty_bar[stage]);
          if (kb == 5) tc::umma_commit(&tmem_full[as]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const float kLog2e = 1.4426950408889634f;
    int it = 0;
    for (int w = blockIdx.x; w < num_tiles; w += gridDim.x, ++it) {
      const int bh = w / per_bh, rem = w - bh * per_bh;
      const int m0 = (rem / p.n_tiles) * BM, n0 = (rem % p.n_tiles) * BN;
      const int row = m0 + q * 32 + lane;
      const bool valid = row < p.Nq;
      const int64_t R = (int64_t)bh * p.Nq + (valid ? row : 0);
      const float nlse = valid ? -p.lse[R] * kLog2e : 0.f;
      const float delta = valid ? p.ws_delta[R] : 0.f;
      const int as = it & 1;
      tc::mbar_wait_hot(&tmem_full[as], (it >> 1) & 1);
      tc::fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 2 * BN);
#pragma unroll 1
      for (int ch = half; ch < BN / 32; ch += 2) {
        float sv[32], dv[32];
        tc::tmem_ld32(taddr + ch * 32, sv);
        tc::tmem_ld32(taddr + BN + ch * 32, dv);
        tc::tmem_ld_wait();
        if (ch + 2 >= BN / 32) {  // this warp's last chunk is in registers: release the accumulator pair
          tc::fence_before_sync();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&tmem_empty[as]);
        }
        const int cbase = n0 + ch * 32;
        if (cbase >= p.Nkp) continue;
        uint32_t pk[16], dk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float pv[2], ds[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            float x = tc::ex2_approx(fmaf(sv[j + u], p.sc, nlse));
            if (cbase + j + u >= p.Nk) x = 0.f;
            pv[u] = x;
            ds[u] = x * (dv[j + u] - delta);
          }
          pk[j >> 1] = pack2(pv[0], pv[1]);
          dk[j >> 1] = pack2(ds[0], ds[1]);
        }
        if (p.stage_out) {
          // thread = row in TMEM, but a row's 64 bytes per tensor are what is contiguous in memory: transpose through the
          // warp's staging tiles so that one store instruction covers 8 rows x 64 B (16 full sectors) instead of 32 rows
          // x 16 B (32 half sectors)
          unsigned char* sp = ostg + (warp - 2) * (2 * 32 * OPITCH);
          unsigned char* sd = sp + 32 * OPITCH;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            *reinterpret_cast<uint4*>(sp + lane * OPITCH + u * 16) = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
            *reinterpret_cast<uint4*>(sd + lane * OPITCH + u * 16) = make_uint4(dk[4 * u], dk[4 * u + 1], dk[4 * u + 2], dk[4 * u + 3]);
          }
          __syncwarp();
          const int piece = lane & 3;
          const bool col_ok = cbase + 8 * piece < p.Nkp;
#pragma unroll
          for (int rr = 0; rr < 32; rr += 8) {
            const int r = rr + (lane >> 2);
            const int grow = m0 + q * 32 + r;
            if (grow < p.Nq && col_ok) {
              const int64_t off = ((int64_t)bh * p.Nq + grow) * p.Nkp + cbase + 8 * piece;
              *reinterpret_cast<uint4*>(p.P + off) = *reinterpret_cast<const uint4*>(sp + r * OPITCH + piece * 16);
              *reinterpret_cast<uint4*>(p.dS + off) = *reinterpret_cast<const uint4*>(sd + r * OPITCH + piece * 16);
            }
          }
          __syncwarp();
        } else if (valid) {
          bf16* prow = p.P + R * p.Nkp + cbase;
          bf16* drow = p.dS + R * p.Nkp + cbase;
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (cbase + 8 * u < p.Nkp) {
              *reinterpret_cast<uint4*>(prow + 8 * u) = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
              *reinterpret_cast<uint4*>(drow + 8 * u) = make_uint4(dk[4 * u], dk[4 * u + 1], dk[4 * u + 2], dk[4 * u + 3]);
            }
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

size_t smem_bytes(int stage_out) {
  return (size_t)STAGES * STAGE_BYTES + 256 + 1024 + 64 + (stage_out ? (size_t)EPI_WARPS * 2 * 32 * OPITCH + 32 : 0);
}

}  // namespace

int svit_attn_bwd_sdp_supported(const svit_attn_args* a) {
  if (a->dtype != SVIT_BF16 || !a->ws_p || !a->ws_ds || !a->ws_e || !a->ws_delta || !a->lse || !a->sel_bwd) return 0;
  const int ne = a->kh + a->kw + a->kt;
  if (a->nep < ne || a->nep % 16 || a->nep > 64) return 0;  // hi | lo halves of E' are whole 16-column MMA steps
  const int64_t Nk = 1 + (int64_t)a->kt * a->kh * a->kw + a->O;
  const int64_t Nq = 1 + (int64_t)a->qt * a->qh * a->qw + a->O;
  if (Nk > MAX_KEYS_PADDED || Nq >= (1ll << 30)) return 0;
  const int64_t tiles = ((Nq + BM - 1) / BM) * ((Nk + BN - 1) / BN) * a->B * a->h;
  if (tiles >= (1ll << 31)) return 0;
  if ((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) | reinterpret_cast<uintptr_t>(a->v) |
       reinterpret_cast<uintptr_t>(a->dout) | reinterpret_cast<uintptr_t>(a->ws_p) | reinterpret_cast<uintptr_t>(a->ws_ds) |
       reinterpret_cast<uintptr_t>(a->ws_e) | reinterpret_cast<uintptr_t>(a->sel_bwd)) & 15)
    return 0;
  return 1;
}

int svit_attn_bwd_sdp(const svit_attn_args* a, cudaStream_t st) {
  Params p;
  p.B = a->B; p.h = a->h;
  p.Lq = a->qt * a->qh * a->qw; p.Lk = a->kt * a->kh * a->kw;
  p.Nq = 1 + p.Lq + a->O; p.Nk = 1 + p.Lk + a->O;
  p.Nkp = (p.Nk + 7) / 8 * 8;
  p.kh = a->kh; p.kw = a->kw; p.kt = a->kt; p.nep = a->nep;
  p.m_tiles = (p.Nq + BM - 1) / BM; p.n_tiles = (p.Nk + BN - 1) / BN;
  p.sc = a->scale * 1.4426950408889634f;
  p.lse = a->lse; p.ws_delta = a->ws_delta;
  p.P = (bf16*)a->ws_p; p.dS = (bf16*)a->ws_ds;
  const uint64_t BH = (uint64_t)a->B * a->h;
  CUtensorMap tq, tk, tdo, tv, te, tsel;
  int rc;
  if ((rc = svit_make_tmap_4d(&tq, a->q, BH, 1, (uint64_t)p.Nq, HD, HD, 0, (uint64_t)p.Nq * HD, BM))) return rc;
  if ((rc = svit_make_tmap_4d(&tk, a->k, BH, 1, (uint64_t)p.Nk, HD, HD, 0, (uint64_t)p.Nk * HD, BN))) return rc;
  if ((rc = svit_make_tmap_4d(&tv, a->v, BH, 1, (uint64_t)p.Nk, HD, HD, 0, (uint64_t)p.Nk * HD, BN))) return rc;
  if ((rc = svit_make_tmap_4d(&tdo, a->dout, (uint64_t)a->B, (uint64_t)a->h, (uint64_t)p.Nq, HD, (uint64_t)a->h * HD, HD,
                              (uint64_t)p.Nq * a->h * HD, BM)))
    return rc;
  // E' = E / scale as bf16 [B h, Nq, hi (nep) | lo (nep)] in the ws_e scratch (the bytes of its fp32 rows;
  // attn_bwd_prep_kernel, e16 mode); Sel = sel_bwd [Nk, nep]
  if ((rc = svit_make_tmap_4d(&te, a->ws_e, BH, 1, (uint64_t)p.Nq, (uint64_t)2 * p.nep, (uint64_t)2 * p.nep, 0,
                              (uint64_t)p.Nq * 2 * p.nep, BM)))
    return rc;
  if ((rc = svit_make_tmap_4d(&tsel, a->sel_bwd, 1, 1, (uint64_t)p.Nk, (uint64_t)p.nep, (uint64_t)p.nep, 0,
                              (uint64_t)p.Nk * p.nep, BN)))
    return rc;
  p.stage_out = getenv("SVIT_SDP_DIRECT_STORES") ? 0 : 1;  // A/B switch
  const size_t smem = smem_bytes(p.stage_out);
  static SvitDevOnce configured;
  if (configured.need(smem)) {
    SVIT_CUDA(cudaFuncSetAttribute(attn_bwd_sdp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured.done(smem);
  }
  const int64_t tiles = (int64_t)p.m_tiles * p.n_tiles * BH;
  const int grid = (int)(tiles < svit_num_sms() ? tiles : svit_num_sms());
  attn_bwd_sdp_kernel<<<grid, NUM_THREADS, smem, st>>>(tq, tk, tdo, tv, te, tsel, p);
  SVIT_CHECK_LAUNCH();
  return 0;
}
