// bf16 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands fed by TMA),
// with the fused epilogue of the SViT hot path.  Same contract as gemm_simt.cu (svit_gemm_args):
//   C[M,N] = residual + sample_scale[row/rps] * ( act(op(A).op(B) + bias) * gelu'(gelu_pre) )
//
// Persistent, warp-specialised, one CTA per SM (320 threads):
//   warp 0    TMA producer: 4-stage ring of {A 128x64, B BNx64} bf16 tiles, 128-byte swizzle
//   warp 1    MMA issuer (one elected lane): tcgen05.mma 128 x BN x 16, fp32 accumulators in TMEM,
//             two accumulator buffers so the MMAs of tile i+1 overlap the epilogue of tile i
//   warps 2-9 epilogue (two per TMEM lane quarter, alternate 32-column chunks): tcgen05.ld -> fp32 staging in
//             smem -> coalesced 16-byte row segments: bias, GELU (A&S erf), gelu' multiply, DropPath scale,
//             residual add, store (bf16 or fp32)
// Operands may be K-major (row-major [rows, K]) or MN-major (row-major [K, rows]) -- the latter is what the
// dgrad / wgrad GEMMs need -- selected through the UMMA descriptors; no transposed copies are made.
#include <mutex>

#include "tc_common.cuh"
#include "../../include/svit_b200.h"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int STAGES = 4;
constexpr int NUM_THREADS = 320;        // TMA warp, MMA warp, 8 epilogue warps
constexpr int EPI_WARPS = 8;
constexpr int CH = 32;                  // accumulator columns per epilogue chunk
constexpr int STG_PITCH = CH * 4 + 16;  // fp32 staging row, padded: conflict-free 16-byte accesses

template <int BN>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STG_BYTES = EPI_WARPS * 32 * STG_PITCH;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES + STG_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;  // + alignment slack
};

struct EpiArgs {
  const float* bias;
  const bf16* residual;
  int64_t ldr;
  const float* sample_scale;
  int64_t rps;
  const bf16* gelu_pre;
  int64_t ldg;
  bf16* pre_out;
  int64_t ldp;
  int act;
  int64_t rows_in, rows_out, row_off;
  void* C;
  int64_t ldc;
  int out_f32;
  int splits;  // split-K: > 1 -> fp32 partial sums are accumulated with atomics into a zeroed C
  int batch, b_inner;   // batched GEMM: work item = (batch, K split, output tile)
  int a_inner;          // A's two-level batch split (1 = flat)
  int64_t strideC;      // elements between consecutive problems' outputs
  float alpha;          // accumulator scale (applied before the bias)
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below bf16 resolution): one RCP, one EX2, 5 FMAs.
// Returns Phi(x) (the GELU gate) and, through `pdf_x`, x * phi(x) for the derivative.
__device__ __forceinline__ float gelu_gate(float x, float& pdf_x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float ex = exp2f(-1.4426950408889634f * z * z);  // exp(-x^2/2)
  const float erfc_half = 0.5f * poly * t * ex;          // 0.5 * erfc(|x|/sqrt2)
  pdf_x = x * 0.3989422804014327f * ex;
  return x >= 0.f ? 1.0f - erfc_half : erfc_half;
}

// Forward GELU for bf16 outputs: x * Phi(x) with Phi(x) = 0.5 (1 + tanh(x (a + b x^2 + c x^4))), coefficients fitted to
// the exact erf form (max |x Phi - gelu_erf| = 2.5e-5 on [-9, 9], x^2 clamped beyond) and MUFU.TANH (rel. err 2^-11):
// total error <= 2.5e-4 |x|, 16x below the bf16 rounding of the stored result.  One MUFU + 7 FP32 ops per element
// (the erf form above needs two MUFUs and is kept for the derivative, where exp(-x^2/2) is needed anyway).
__device__ __forceinline__ float gelu_fwd_fast(float x) {
  const float x2 = fminf(x * x, 81.0f);
  const float u = x * fmaf(x2, fmaf(x2, -3.51516789e-04f, 3.70056460e-02f), 7.97507884e-01f);
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(u));
  const float hx = 0.5f * x;
  return fmaf(hx, th, hx);
}

// Coalesced write-out of one staged chunk (32 rows x 32 fp32 columns) owned by this warp: lanes 4r..4r+3 cover
// the 32 columns of row r in 8-column units; bias, GELU, gelu' multiply, DropPath scale and residual happen here.
__device__ __forceinline__ void store_chunk(const EpiArgs& e, const unsigned char* stg, int lane, int64_t m_base,
                                            int64_t n_base, int64_t M, int64_t N) {
  const int c8 = (lane & 3) * 8;
  const int64_t n = n_base + c8;
  if (n >= N) return;  // N % 8 != 0: the last unit also writes the zero pad columns up to the next multiple of 8
  float bias[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (e.bias) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(e.bias + n));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(e.bias + n + 4));
    bias[0] = b0.x; bias[1] = b0.y; bias[2] = b0.z; bias[3] = b0.w;
    bias[4] = b1.x; bias[5] = b1.y; bias[6] = b1.z; bias[7] = b1.w;
  }
  // issue every global read of this chunk first (residual / gelu_pre rows): their DRAM latency then overlaps
  // the shared-memory reads and the math of all four row groups instead of serialising per row
  uint4 res[4], gpre[4];
  int64_t orow[4];
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int64_t m = m_base + it * 8 + (lane >> 2);
    const int64_t mc = m < M ? m : M - 1;
    orow[it] = e.rows_in > 0 ? (mc / e.rows_in) * e.rows_out + e.row_off + (mc % e.rows_in) : mc;
    if (e.residual) res[it] = *reinterpret_cast<const uint4*>(e.residual + orow[it] * e.ldr + n);
    if (e.gelu_pre) gpre[it] = *reinterpret_cast<const uint4*>(e.gelu_pre + mc * e.ldg + n);
  }
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int r = it * 8 + (lane >> 2);
    const int64_t m = m_base + r;
    if (m >= M) continue;
    const float4 lo = *reinterpret_cast<const float4*>(stg + r * STG_PITCH + c8 * 4);
    const float4 hi = *reinterpret_cast<const float4*>(stg + r * STG_PITCH + c8 * 4 + 16);
    float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], e.alpha, bias[i]);
    if (e.pre_out) {
      uint4 o = {pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7])};
      *reinterpret_cast<uint4*>(e.pre_out + m * e.ldp + n) = o;
    }
    if (e.act == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = gelu_fwd_fast(v[i]);
    }
    if (e.gelu_pre) {
      const __nv_bfloat162* gp = reinterpret_cast<const __nv_bfloat162*>(&gpre[it]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(gp[i]);
        float px, py;
        const float gx = gelu_gate(f.x, px), gy = gelu_gate(f.y, py);
        v[2 * i] *= gx + px;
        v[2 * i + 1] *= gy + py;
      }
    }
    if (e.sample_scale) {
      const float sc = e.sample_scale[m / e.rps];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] *= sc;
    }
    if (e.residual) {
      const __nv_bfloat162* gp = reinterpret_cast<const __nv_bfloat162*>(&res[it]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(gp[i]);
        v[2 * i] += f.x;
        v[2 * i + 1] += f.y;
      }
    }
    if (e.out_f32) {
      float* dst = reinterpret_cast<float*>(e.C) + orow[it] * e.ldc + n;
      if (e.splits > 1) {
        atomicAdd(reinterpret_cast<float4*>(dst), make_float4(v[0], v[1], v[2], v[3]));
        atomicAdd(reinterpret_cast<float4*>(dst + 4), make_float4(v[4], v[5], v[6], v[7]));
      } else {
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
    } else {
      uint4 o = {pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7])};
      *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(e.C) + orow[it] * e.ldc + n) = o;
    }
  }
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, int64_t M,
               int64_t N, int64_t K, EpiArgs e) {
  using L = SmemLayout<BN>;
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment by pointer offset (not an integer round trip) so accesses stay in the shared state space
  unsigned char* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* stg_base = smem + STAGES * L::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m_tiles = (M + BM - 1) / BM, n_tiles = (N + BN - 1) / BN;
  const int64_t out_tiles = m_tiles * n_tiles;
  const int64_t per_batch = out_tiles * e.splits;           // work item = (batch, K split, output tile)
  const int64_t num_tiles = per_batch * e.batch;
  const int num_kb_all = (int)((K + BK - 1) / BK);
  const int kb_per = (num_kb_all + e.splits - 1) / e.splits;  // K blocks per split
  constexpr uint32_t TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_a);
    tc::prefetch_tmap(&tmap_b);
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
      for (int64_t w = blockIdx.x; w < num_tiles; w += gridDim.x) {
        const int bi = (int)(w / per_batch);
        const int64_t wb = w - (int64_t)bi * per_batch;
        const int64_t t = wb % out_tiles;
        const int kb0 = (int)(wb / out_tiles) * kb_per;
        const int kb1 = kb0 + kb_per < num_kb_all ? kb0 + kb_per : num_kb_all;
        const int m0 = (int)((t / n_tiles) * BM), n0 = (int)((t % n_tiles) * BN);
        const int bo = bi / e.b_inner, bn = bi - bo * e.b_inner;  // B's two-level batch coordinate
        const int ao = bi / e.a_inner, an = bi - ao * e.a_inner;  // A's
        for (int kb = kb0; kb < kb1; ++kb) {
          tc::mbar_wait(&empty_bar[stage], phase ^ 1);
          unsigned char* sa = smem + stage * L::STAGE_BYTES;
          unsigned char* sb = sa + L::A_BYTES;
          tc::mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
          if (!A_MN) {
            tc::tma_load_4d(sa, &tmap_a, &full_bar[stage], kb * BK, m0, an, ao);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tc::tma_load_4d(sa + j * 8192, &tmap_a, &full_bar[stage], m0 + 64 * j, kb * BK, an, ao);
          }
          if (!B_MN) {
            tc::tma_load_4d(sb, &tmap_b, &full_bar[stage], kb * BK, n0, bn, bo);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tc::tma_load_4d(sb + j * 8192, &tmap_b, &full_bar[stage], n0 + 64 * j, kb * BK, bn, bo);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (tc::elect_one()) {
      constexpr uint32_t idesc = tc::idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int64_t w = blockIdx.x; w < num_tiles; w += gridDim.x, ++it) {
        const int kb0 = (int)((w % per_batch) / out_tiles) * kb_per;
        const int kb1 = kb0 + kb_per < num_kb_all ? kb0 + kb_per : num_kb_all;
        const int as = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        tc::mbar_wait_hot(&tmem_empty[as], acc_phase ^ 1);
        tc::fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          tc::mbar_wait_hot(&full_bar[stage], phase);
          tc::fence_after_sync();
          const uint32_t sa = tc::smem_u32(smem + stage * L::STAGE_BYTES);
          const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = A_MN ? tc::smem_desc_sw128(sa + k * 2048, 8192, 1024) : tc::smem_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? tc::smem_desc_sw128(sb + k * 2048, 8192, 1024) : tc::smem_desc_sw128(sb + k * 32, 16, 1024);
            tc::umma_bf16_ss(d_tmem, da, db, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
          }
          tc::umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs have read it
          if (kb == kb1 - 1) tc::umma_commit(&tmem_full[as]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;  // the two warps of a quarter take alternate 32-column chunks
    unsigned char* stg = stg_base + (warp - 2) * 32 * STG_PITCH;
    constexpr int NCH = BN / CH;
    constexpr int LAST_CH0 = ((NCH - 1) & 1) == 0 ? NCH - 1 : NCH - 2;  // last chunk index handled by half 0
    constexpr int LAST_CH1 = ((NCH - 1) & 1) == 1 ? NCH - 1 : NCH - 2;
    int it = 0;
    for (int64_t w = blockIdx.x; w < num_tiles; w += gridDim.x, ++it) {
      const int64_t t = (w % per_batch) % out_tiles;
      EpiArgs eb = e;  // this problem's output
      if (e.batch > 1) {
        const int64_t off = (w / per_batch) * e.strideC;
        eb.C = e.out_f32 ? (void*)(reinterpret_cast<float*>(e.C) + off) : (void*)(reinterpret_cast<bf16*>(e.C) + off);
      }
      const int as = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int64_t m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * BN;
      tc::mbar_wait_hot(&tmem_full[as], acc_phase);
      tc::fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
      const int last = half ? LAST_CH1 : LAST_CH0;
#pragma unroll 1
      for (int ch = half; ch < NCH; ch += 2) {
        float v[CH];
        tc::tmem_ld32(taddr + ch * CH, v);
        tc::tmem_ld_wait();
        if (ch == last) {  // this warp's part of the accumulator is in registers: release the TMEM buffer
          tc::fence_before_sync();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&tmem_empty[as]);
        }
#pragma unroll
        for (int j = 0; j < CH; j += 4)
          *reinterpret_cast<float4*>(stg + lane * STG_PITCH + j * 4) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
        store_chunk(eb, stg, lane, m0 + q * 32, n0 + ch * CH, M, N);
        __syncwarp();
      }
      if (NCH == 1 && half == 1) {  // BN == 32 never instantiated; keeps the arrive count uniform
        if (lane == 0) tc::mbar_arrive(&tmem_empty[as]);
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

template <int BN, bool A_MN, bool B_MN>
int launch(const svit_gemm_args* a, cudaStream_t st) {
  using L = SmemLayout<BN>;
  CUtensorMap ta, tb;
  int rc;
  const uint64_t nb = a->batch > 1 ? (uint64_t)a->batch : 1, bin = (a->batch > 1 && a->b_inner > 1) ? (uint64_t)a->b_inner : 1;
  const uint64_t ain = (a->batch > 1 && a->a_inner > 1) ? (uint64_t)a->a_inner : 1;
  if (!A_MN) rc = svit_make_tmap_4d(&ta, a->A, nb / ain, ain, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda, (uint64_t)a->strideA_inner, (uint64_t)a->strideA, BM);
  else rc = svit_make_tmap_4d(&ta, a->A, nb / ain, ain, (uint64_t)a->K, (uint64_t)a->M, (uint64_t)a->lda, (uint64_t)a->strideA_inner, (uint64_t)a->strideA, BK);
  if (rc) return rc;
  if (!B_MN) rc = svit_make_tmap_4d(&tb, a->B, nb / bin, bin, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldb, (uint64_t)a->strideB_inner, (uint64_t)a->strideB, BN);
  else rc = svit_make_tmap_4d(&tb, a->B, nb / bin, bin, (uint64_t)a->K, (uint64_t)a->N, (uint64_t)a->ldb, (uint64_t)a->strideB_inner, (uint64_t)a->strideB, BK);
  if (rc) return rc;
  EpiArgs e;
  e.bias = a->bias;
  e.residual = (const bf16*)a->residual; e.ldr = a->ldr;
  e.sample_scale = a->sample_scale; e.rps = a->rows_per_sample;
  e.gelu_pre = (const bf16*)a->gelu_pre; e.ldg = a->ldg;
  e.pre_out = (bf16*)a->pre_out; e.ldp = a->ldp;
  e.act = a->act;
  e.rows_in = a->rows_in; e.rows_out = a->rows_out; e.row_off = a->row_off;
  e.C = a->C; e.ldc = a->ldc;
  e.out_f32 = a->out_dtype == SVIT_F32;
  // split-K for the weight-gradient shape (few output tiles, very long reduction): fp32 partials via atomics
  e.splits = 1;
  e.batch = (int)nb; e.b_inner = (int)bin; e.a_inner = (int)ain; e.strideC = a->strideC;
  e.alpha = a->alpha == 0.f ? 1.f : a->alpha;
  if (nb == 1) {
    const int64_t tiles0 = ((a->M + BM - 1) / BM) * ((a->N + BN - 1) / BN);
    const int64_t nkb = (a->K + BK - 1) / BK;
    const bool plain = e.out_f32 && !a->bias && !a->residual && !a->gelu_pre && !a->pre_out && !a->sample_scale &&
                       a->act == 0 && a->rows_in == 0 && a->ldc == a->N;
    if (plain && tiles0 * 2 <= svit_num_sms() && nkb >= 16) {
      int64_t sp = svit_num_sms() / tiles0;
      if (sp > nkb / 4) sp = nkb / 4;
      if (sp > 1) {
        const int64_t per = (nkb + sp - 1) / sp;
        e.splits = (int)((nkb + per - 1) / per);  // no empty split
        SVIT_CUDA(cudaMemsetAsync(a->C, 0, (size_t)a->M * a->N * sizeof(float), st));
      }
    }
  }
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN>;
  static SvitDevOnce configured;
  if (configured.need()) {
    SVIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured.done();
  }
  const int64_t tiles = ((a->M + BM - 1) / BM) * ((a->N + BN - 1) / BN) * e.splits * e.batch;
  const int grid = (int)(tiles < svit_num_sms() ? tiles : svit_num_sms());
  kern<<<grid, NUM_THREADS, L::TOTAL, st>>>(ta, tb, a->M, a->N, a->K, e);
  SVIT_CHECK_LAUNCH();
  return 0;
}

template <bool A_MN, bool B_MN>
int dispatch_bn(const svit_gemm_args* a, cudaStream_t st) {
  const int64_t N = a->N;
  if (!B_MN) {
    if (N % 192 == 0 || N > 384) return launch<192, A_MN, B_MN>(a, st);
    if (N % 96 == 0) return launch<96, A_MN, B_MN>(a, st);
    if (N > 128) return launch<192, A_MN, B_MN>(a, st);
    return N > 64 ? launch<128, A_MN, B_MN>(a, st) : launch<64, A_MN, B_MN>(a, st);
  }
  // MN-major B tiles are built from 64-wide swizzle atoms
  if (N % 192 == 0 || N > 256) return launch<192, A_MN, B_MN>(a, st);
  if (N > 64) return launch<128, A_MN, B_MN>(a, st);
  return launch<64, A_MN, B_MN>(a, st);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

svit_tmap_encode_fn svit_get_tmap_encode() {
  static svit_tmap_encode_fn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<svit_tmap_encode_fn>(p);
  });
  return fn;
}

int svit_gemm_tc_supported(const svit_gemm_args* a) {
  if (a->dtype != SVIT_BF16) return 0;
  if (a->batch > 1) {  // batched form: no epilogue extras except bias, 16-byte aligned problem strides
    if (a->residual || a->gelu_pre || a->pre_out || a->sample_scale || a->act || a->rows_in) return 0;
    if (a->strideA % 8 || a->strideB % 8 || a->strideC % 8 || (a->b_inner > 1 && (a->strideB_inner % 8 || a->batch % a->b_inner))) return 0;
    if (a->a_inner > 1 && (a->strideA_inner % 8 || a->batch % a->a_inner)) return 0;
  }
  if (a->out_dtype != SVIT_BF16 && a->out_dtype != SVIT_F32) return 0;
  if (a->K < 8 || a->N < 8 || a->M < 1) return 0;
  if (a->lda % 8 || a->ldb % 8) return 0;
  // ragged N (e.g. the 457 keys of a score matrix): the last 8-column unit of a row also writes zeros into the pad
  // columns, so the row pitch must cover them; no bias vector to read past
  if (a->N % 8 && (a->bias || a->ldc < ((a->N + 7) / 8) * 8 || a->rows_in || a->residual || a->gelu_pre || a->pre_out)) return 0;
  if (a->out_dtype == SVIT_BF16 ? (a->ldc % 8) : (a->ldc % 4)) return 0;
  if (!aligned16(a->A) || !aligned16(a->B) || !aligned16(a->C)) return 0;
  if (a->residual && (!aligned16(a->residual) || a->ldr % 8)) return 0;
  if (a->gelu_pre && (!aligned16(a->gelu_pre) || a->ldg % 8)) return 0;
  if (a->pre_out && (!aligned16(a->pre_out) || a->ldp % 8)) return 0;
  if (a->M >= (1ll << 31) || a->N >= (1ll << 31) || a->K >= (1ll << 31)) return 0;
  // TMA needs 16-byte multiples for the pitches (checked above) only: ragged extents are zero-filled by the copy
  return 1;
}

int svit_gemm_tc_tma_supported(const svit_gemm_args* a);  // gemm_tc2.cu
int svit_gemm_tc_tma(const svit_gemm_args* a, cudaStream_t st);

int svit_gemm_tc(const svit_gemm_args* a, cudaStream_t st) {
  const bool plain_scale = a->alpha == 0.f || a->alpha == 1.f;
  if (a->batch <= 1 && plain_scale && a->N % 8 == 0 && a->K % 8 == 0 && svit_gemm_tc_tma_supported(a)) return svit_gemm_tc_tma(a, st);
  if (a->ln_stats) return SVIT_ENOTSUP;  // the folded-LayerNorm epilogue exists in the TMA-store kernel only
  const bool a_mn = a->transA != 0;  // A stored [K, M]
  const bool b_mn = a->transB == 0;  // B stored [K, N]
  if (!a_mn && !b_mn) return dispatch_bn<false, false>(a, st);
  if (!a_mn && b_mn) return dispatch_bn<false, true>(a, st);
  if (a_mn && b_mn) return dispatch_bn<true, true>(a, st);
  return dispatch_bn<true, false>(a, st);
}

extern "C" int svit_destroy(void) { return 0; }
