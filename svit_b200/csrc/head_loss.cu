// the step right after the hot path, as two kernels.
//
// svit_head_fwd   SViTHead.forward in eval mode (video_model_builder.py:507-546) in ONE launch: class projection +
//                 softmax / sigmoid, box MLP + sigmoid, box-score linear (+ sigmoid), contact-state linear on the two
//                 hand slots (+ softmax), the (score | box) concatenation and the fp32 object descriptors.  The reference
//                 (and round 1 here) issues four skinny GEMMs plus sigmoid / softmax / cat kernels on [B, 65, 768] rows.
// svit_haog_loss  VideoImageLoss._haog_loss (models/losses.py:50-92, 138-155; GIoU of utils/box_ops.py:41-77): box L1,
//                 box-score BCE, GIoU and contact-state cross entropy as masked means over paired boxes, values AND the
//                 gradients w.r.t. the predictions, in one launch without the reference's `mask.sum() > 0` host round
//                 trips and without the N x N GIoU matrix.  Inputs are tiny ([B, T, O, 5]): one CTA, two passes.
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ head
// One CTA per row of x [B, 1 + T*O, C]; a warp per output neuron (lanes stride the C channels), fp32 accumulation.
// Row 0 of a sample = cls: 174 (NC) class logits, then softmax / sigmoid in place.  Row 1 + t*O + o = object token:
// 4 box coordinates + 1 score (+ 5 contact states for o < 2) and the fp32 copy of the token.
template <typename T>
__global__ void __launch_bounds__(256) head_fwd_kernel(const T* __restrict__ x, const float* __restrict__ wp,
                                                       const float* __restrict__ bp, const float* __restrict__ wb,
                                                       const float* __restrict__ bb, const float* __restrict__ ws,
                                                       const float* __restrict__ bs, const float* __restrict__ wc,
                                                       const float* __restrict__ bc, float* __restrict__ logits,
                                                       float* __restrict__ probs, float* __restrict__ obj_desc,
                                                       float* __restrict__ pred_bboxes, float* __restrict__ pred_contact,
                                                       int B, int Tx, int O, int C, int NC, int act_sigmoid, int eval_mode) {
  extern __shared__ float sm[];  // [C] the row in fp32, then [NC] logits (cls row)
  float* row = sm;
  float* lg = sm + C;
  const int rows_per = 1 + Tx * O;
  const int b = blockIdx.x / rows_per, r = blockIdx.x - b * rows_per;
  const T* xr = x + ((int64_t)b * rows_per + r) * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) row[c] = to_f(xr[c]);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  auto dot = [&](const float* __restrict__ w) {
    float a = 0.f;
    for (int c = lane; c < C; c += 32) a = fmaf(row[c], __ldg(w + c), a);
    return warp_sum(a);
  };
  if (r == 0) {
    for (int n = warp; n < NC; n += nwarps) {
      const float v = dot(wp + (int64_t)n * C) + bp[n];
      if (lane == 0) lg[n] = v;
    }
    __syncthreads();
    for (int n = threadIdx.x; n < NC; n += blockDim.x) logits[(int64_t)b * NC + n] = lg[n];
    if (!probs) return;
    if (act_sigmoid) {
      for (int n = threadIdx.x; n < NC; n += blockDim.x) probs[(int64_t)b * NC + n] = 1.f / (1.f + __expf(-lg[n]));
      return;
    }
    __shared__ float red[8];
    float m = -INFINITY;
    for (int n = threadIdx.x; n < NC; n += blockDim.x) m = fmaxf(m, lg[n]);
    m = warp_max(m);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
    for (int i = 1; i < nwarps; ++i) m = fmaxf(m, red[i]);
    __syncthreads();
    float s = 0.f;
    for (int n = threadIdx.x; n < NC; n += blockDim.x) s += expf(lg[n] - m);
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    s = 0.f;
    for (int i = 0; i < nwarps; ++i) s += red[i];
    for (int n = threadIdx.x; n < NC; n += blockDim.x) probs[(int64_t)b * NC + n] = expf(lg[n] - m) / s;
    return;
  }
  const int to = r - 1, t = to / O, o = to - t * O;
  float* od = obj_desc + ((int64_t)b * Tx * O + to) * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) od[c] = row[c];
  // outputs 0..3: box (sigmoid), 4: score, 5..9: contact (slots 0, 1 only)
  const int nout = o < 2 ? 10 : 5;
  for (int n = warp; n < nout; n += nwarps) {
    float v;
    if (n < 4) v = dot(wb + (int64_t)n * C) + bb[n];
    else if (n == 4) v = dot(ws) + bs[0];
    else v = dot(wc + (int64_t)(n - 5) * C) + bc[n - 5];
    if (lane == 0) lg[n] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float* pb = pred_bboxes + ((int64_t)b * Tx * O + to) * 5;
    pb[0] = eval_mode ? 1.f / (1.f + expf(-lg[4])) : lg[4];
    for (int i = 0; i < 4; ++i) pb[1 + i] = 1.f / (1.f + expf(-lg[i]));
    if (o < 2) {
      float* pc = pred_contact + (((int64_t)b * Tx + t) * 2 + o) * 5;
      if (eval_mode) {
        float m = lg[5];
        for (int i = 1; i < 5; ++i) m = fmaxf(m, lg[5 + i]);
        float s = 0.f;
        for (int i = 0; i < 5; ++i) s += expf(lg[5 + i] - m);
        for (int i = 0; i < 5; ++i) pc[i] = expf(lg[5 + i] - m) / s;
      } else {
        for (int i = 0; i < 5; ++i) pc[i] = lg[5 + i];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ losses
struct Pair {
  float loss_giou;
  float g[4];  // d(1 - giou) / d(cx, cy, w, h) of the prediction
};

// torch.minimum / maximum route the gradient to the smaller / larger argument and split it on ties
__device__ __forceinline__ float pick_lt(float a, float b) { return a < b ? 1.f : (a == b ? 0.5f : 0.f); }
__device__ __forceinline__ float pick_gt(float a, float b) { return a > b ? 1.f : (a == b ? 0.5f : 0.f); }

__device__ __forceinline__ Pair giou_pair(const float* s, const float* t) {
  // cxcywh -> xyxy (utils/box_ops.py:26-30)
  const float ax0 = s[0] - 0.5f * s[2], ay0 = s[1] - 0.5f * s[3], ax1 = s[0] + 0.5f * s[2], ay1 = s[1] + 0.5f * s[3];
  const float bx0 = t[0] - 0.5f * t[2], by0 = t[1] - 0.5f * t[3], bx1 = t[0] + 0.5f * t[2], by1 = t[1] + 0.5f * t[3];
  const float aw = ax1 - ax0, ah = ay1 - ay0;
  const float area1 = aw * ah, area2 = (bx1 - bx0) * (by1 - by0);
  const float iw_raw = fminf(ax1, bx1) - fmaxf(ax0, bx0), ih_raw = fminf(ay1, by1) - fmaxf(ay0, by0);
  const float iw = fmaxf(iw_raw, 0.f), ih = fmaxf(ih_raw, 0.f);
  const float inter = iw * ih, uni = area1 + area2 - inter, iou = inter / uni;
  const float cw_raw = fmaxf(ax1, bx1) - fminf(ax0, bx0), ch_raw = fmaxf(ay1, by1) - fminf(ay0, by0);
  const float cw = fmaxf(cw_raw, 0.f), ch = fmaxf(ch_raw, 0.f);
  const float areac = cw * ch;
  Pair p;
  p.loss_giou = 1.f - (iou - (areac - uni) / areac);
  // reverse mode.  L = 1 - iou + (areac - uni) / areac = 2 - inter/uni - uni/areac
  const float d_inter0 = -1.f / uni;                               // dL/d inter through iou (uni held)
  const float d_uni = inter / (uni * uni) - 1.f / areac;           // dL/d uni
  const float d_areac = uni / (areac * areac);                     // dL/d areac
  const float d_inter = d_inter0 - d_uni;                          // uni = area1 + area2 - inter
  const float d_area1 = d_uni;
  const float d_iw = (iw_raw >= 0.f ? 1.f : 0.f) * d_inter * ih, d_ih = (ih_raw >= 0.f ? 1.f : 0.f) * d_inter * iw;
  const float d_cw = (cw_raw >= 0.f ? 1.f : 0.f) * d_areac * ch, d_ch = (ch_raw >= 0.f ? 1.f : 0.f) * d_areac * cw;
  // xyxy gradients of the prediction box
  const float gx1 = d_iw * pick_lt(ax1, bx1) + d_cw * pick_gt(ax1, bx1) + d_area1 * ah;
  const float gx0 = -d_iw * pick_gt(ax0, bx0) - d_cw * pick_lt(ax0, bx0) - d_area1 * ah;
  const float gy1 = d_ih * pick_lt(ay1, by1) + d_ch * pick_gt(ay1, by1) + d_area1 * aw;
  const float gy0 = -d_ih * pick_gt(ay0, by0) - d_ch * pick_lt(ay0, by0) - d_area1 * aw;
  p.g[0] = gx0 + gx1;
  p.g[1] = gy0 + gy1;
  p.g[2] = 0.5f * (gx1 - gx0);
  p.g[3] = 0.5f * (gy1 - gy0);
  return p;
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < nw; ++i) s += red[i];
  return s;
}

// pred [N, 5] = (score logit, cx, cy, w, h); tar [N, tc] with tc = 4 (all-zero row = no box) or 5 (leading soft mask);
// contact [M, 5] logits, ctar [M] int64 (-1 = not annotated).
// out[0..3] = (l1, bce, giou, contact ce); d_l1 / d_giou [N, 4], d_bce [N], d_ce [M, 5] = gradients of the four terms.
__global__ void __launch_bounds__(1024) haog_loss_kernel(const float* __restrict__ pred, const float* __restrict__ tar, int tc,
                                                         int64_t N, const float* __restrict__ contact,
                                                         const int64_t* __restrict__ ctar, int64_t M, float* __restrict__ out,
                                                         float* __restrict__ d_l1, float* __restrict__ d_bce,
                                                         float* __restrict__ d_giou, float* __restrict__ d_ce) {
  __shared__ float red[32];
  float s_l1 = 0.f, s_bce = 0.f, s_giou = 0.f, s_ce = 0.f, n_box = 0.f, n_ct = 0.f;
  for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
    const float* p = pred + i * 5;
    const float* t = tar + i * tc + (tc == 5 ? 1 : 0);
    const float mcont = tc == 5 ? tar[i * 5] : ((t[0] == 0.f && t[1] == 0.f && t[2] == 0.f && t[3] == 0.f) ? 0.f : 1.f);
    const bool m = tc == 5 ? mcont > 0.5f : mcont != 0.f;
    const float x = p[0];
    s_bce += fmaxf(x, 0.f) - x * mcont + log1pf(expf(-fabsf(x)));  // binary_cross_entropy_with_logits
    if (m) {
      n_box += 1.f;
      s_l1 += fabsf(p[1] - t[0]) + fabsf(p[2] - t[1]) + fabsf(p[3] - t[2]) + fabsf(p[4] - t[3]);
      s_giou += giou_pair(p + 1, t).loss_giou;
    }
  }
  for (int64_t i = threadIdx.x; i < M; i += blockDim.x) {
    const int64_t y = ctar[i];
    if (y < 0) continue;
    const float* c = contact + i * 5;
    float mx = c[0];
    for (int k = 1; k < 5; ++k) mx = fmaxf(mx, c[k]);
    float se = 0.f;
    for (int k = 0; k < 5; ++k) se += expf(c[k] - mx);
    s_ce += mx + logf(se) - c[y];
    n_ct += 1.f;
  }
  s_l1 = block_sum(s_l1, red);
  s_bce = block_sum(s_bce, red);
  s_giou = block_sum(s_giou, red);
  s_ce = block_sum(s_ce, red);
  n_box = block_sum(n_box, red);
  n_ct = block_sum(n_ct, red);
  const float inv_box = 1.f / fmaxf(n_box, 1.f), inv_ct = 1.f / fmaxf(n_ct, 1.f), inv_n = N > 0 ? 1.f / (float)N : 0.f;
  if (threadIdx.x == 0) {
    out[0] = s_l1 * inv_box * 0.25f;
    out[1] = s_bce * inv_n;
    out[2] = s_giou * inv_box;
    out[3] = s_ce * inv_ct;
  }
  for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
    const float* p = pred + i * 5;
    const float* t = tar + i * tc + (tc == 5 ? 1 : 0);
    const float mcont = tc == 5 ? tar[i * 5] : ((t[0] == 0.f && t[1] == 0.f && t[2] == 0.f && t[3] == 0.f) ? 0.f : 1.f);
    const bool m = tc == 5 ? mcont > 0.5f : mcont != 0.f;
    d_bce[i] = (1.f / (1.f + expf(-p[0])) - mcont) * inv_n;
    if (m) {
      const Pair g = giou_pair(p + 1, t);
      for (int k = 0; k < 4; ++k) {
        const float d = p[1 + k] - t[k];
        d_l1[i * 4 + k] = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * inv_box * 0.25f;
        d_giou[i * 4 + k] = g.g[k] * inv_box;
      }
    } else {
      for (int k = 0; k < 4; ++k) d_l1[i * 4 + k] = d_giou[i * 4 + k] = 0.f;
    }
  }
  for (int64_t i = threadIdx.x; i < M; i += blockDim.x) {
    const int64_t y = ctar[i];
    float* d = d_ce + i * 5;
    if (y < 0) {
      for (int k = 0; k < 5; ++k) d[k] = 0.f;
      continue;
    }
    const float* c = contact + i * 5;
    float mx = c[0];
    for (int k = 1; k < 5; ++k) mx = fmaxf(mx, c[k]);
    float e[5], se = 0.f;
    for (int k = 0; k < 5; ++k) {
      e[k] = expf(c[k] - mx);
      se += e[k];
    }
    for (int k = 0; k < 5; ++k) d[k] = (e[k] / se - (k == y ? 1.f : 0.f)) * inv_ct;
  }
}

}  // namespace

extern "C" {

int svit_head_fwd(const void* x, const float* w_proj, const float* b_proj, const float* w_box, const float* b_box,
                  const float* w_score, const float* b_score, const float* w_contact, const float* b_contact, float* logits,
                  float* probs, float* obj_desc, float* pred_bboxes, float* pred_contact, int B, int Tx, int O, int C,
                  int num_classes, int act_sigmoid, int eval_mode, int dtype, void* stream) {
  if (B < 0 || Tx < 1 || O < 2 || C < 1 || num_classes < 1 || !x || !logits || !obj_desc || !pred_bboxes || !pred_contact)
    return SVIT_EINVAL;
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)(C + (num_classes > 16 ? num_classes : 16)) * sizeof(float);
  if (smem > 48 * 1024) return SVIT_ENOTSUP;
  const unsigned grid = (unsigned)((int64_t)B * (1 + Tx * O));
  if (dtype == SVIT_F32)
    head_fwd_kernel<float><<<grid, 256, smem, st>>>((const float*)x, w_proj, b_proj, w_box, b_box, w_score, b_score, w_contact,
                                                    b_contact, logits, probs, obj_desc, pred_bboxes, pred_contact, B, Tx, O, C,
                                                    num_classes, act_sigmoid, eval_mode);
  else if (dtype == SVIT_BF16)
    head_fwd_kernel<bf16><<<grid, 256, smem, st>>>((const bf16*)x, w_proj, b_proj, w_box, b_box, w_score, b_score, w_contact,
                                                   b_contact, logits, probs, obj_desc, pred_bboxes, pred_contact, B, Tx, O, C,
                                                   num_classes, act_sigmoid, eval_mode);
  else
    return SVIT_EINVAL;
  SVIT_CHECK_LAUNCH();
  return 0;
}

int svit_haog_loss(const float* pred_bboxes, const float* target_boxes, int target_cols, int64_t n_boxes,
                   const float* pred_contact, const int64_t* target_contact, int64_t n_contact, float* losses, float* d_l1,
                   float* d_bce, float* d_giou, float* d_ce, void* stream) {
  if ((target_cols != 4 && target_cols != 5) || n_boxes < 0 || n_contact < 0 || !losses) return SVIT_EINVAL;
  if ((n_boxes > 0 && (!pred_bboxes || !target_boxes || !d_l1 || !d_bce || !d_giou)) ||
      (n_contact > 0 && (!pred_contact || !target_contact || !d_ce)))
    return SVIT_EINVAL;
  haog_loss_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(pred_bboxes, target_boxes, target_cols, n_boxes, pred_contact,
                                                        target_contact, n_contact, losses, d_l1, d_bce, d_giou, d_ce);
  SVIT_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
