// Fused AdamW over every parameter tensor in two launches: the reference steps 405 small tensors through
// torch.optim.AdamW after clip_grad_norm_ (models/optimizer.py:89-104, tools/train_net.py:133-151).  Here a device-side
// table {param, grad, exp_avg, exp_avg_sq, numel, weight_decay} per tensor plus a chunk -> (tensor, offset) map drive
//   (1) svit_grad_sqnorm : sum of squares of all gradients (one atomic per CTA)  -> total norm, no host round trip
//   (2) svit_adamw_step  : clip coefficient from that norm, decoupled weight decay, moment updates, parameter update
// in the operation order of torch's implementation (mul_(1 - lr wd); lerp; mul/addcmul; sqrt / sqrt(bc2) + eps; addcdiv).
#include "common.cuh"
#include "../../include/svit_b200.h"

namespace {

__global__ void __launch_bounds__(256) grad_sqnorm_kernel(const svit_optim_tensor* __restrict__ tab,
                                                          const int32_t* __restrict__ chunk_tensor,
                                                          const int64_t* __restrict__ chunk_start, int nchunks, int chunk,
                                                          float* __restrict__ out) {
  __shared__ float red[8];
  float acc = 0.f;
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const svit_optim_tensor t = tab[chunk_tensor[c]];
    const int64_t s = chunk_start[c];
    const int64_t e = s + chunk < t.numel ? s + chunk : t.numel;
    const float* g = reinterpret_cast<const float*>(t.grad);
    for (int64_t i = s + threadIdx.x; i < e; i += blockDim.x) {
      const float v = g[i];
      acc = fmaf(v, v, acc);
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = red[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffu, v, o);
    if (threadIdx.x == 0) atomicAdd(out, v);
  }
}

__global__ void __launch_bounds__(256) adamw_kernel(const svit_optim_tensor* __restrict__ tab,
                                                    const int32_t* __restrict__ chunk_tensor,
                                                    const int64_t* __restrict__ chunk_start, int nchunks, int chunk, float lr,
                                                    float beta1, float beta2, float eps, float bc1, float sqrt_bc2,
                                                    float max_norm, const float* __restrict__ sqnorm,
                                                    const float* __restrict__ hyper) {
  if (hyper) {  // step-dependent scalars from device memory: the launch can be replayed from a CUDA graph
    lr = hyper[0];
    bc1 = hyper[1];
    sqrt_bc2 = hyper[2];
  }
  float coef = 1.f;
  if (max_norm > 0.f && sqnorm) {  // torch.nn.utils.clip_grad_norm_: max_norm / (total_norm + 1e-6), clamped to 1
    coef = max_norm / (sqrtf(*sqnorm) + 1e-6f);
    coef = coef > 1.f ? 1.f : coef;
  }
  const float step_size = lr / bc1;
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const svit_optim_tensor t = tab[chunk_tensor[c]];
    const int64_t s = chunk_start[c];
    const int64_t e = s + chunk < t.numel ? s + chunk : t.numel;
    float* p = reinterpret_cast<float*>(t.param);
    const float* g = reinterpret_cast<const float*>(t.grad);
    float* m = reinterpret_cast<float*>(t.exp_avg);
    float* v = reinterpret_cast<float*>(t.exp_avg_sq);
    const float decay = 1.f - lr * t.weight_decay;
    for (int64_t i = s + threadIdx.x; i < e; i += blockDim.x) {
      const float gi = g[i] * coef;
      float pi = p[i] * decay;
      float mi = m[i];
      mi = mi + (1.f - beta1) * (gi - mi);  // lerp_(grad, 1 - beta1)
      float vi = v[i] * beta2;
      vi = fmaf((1.f - beta2) * gi, gi, vi);  // addcmul_(grad, grad, value = 1 - beta2)
      const float denom = sqrtf(vi) / sqrt_bc2 + eps;
      pi = pi - step_size * (mi / denom);
      p[i] = pi;
      m[i] = mi;
      v[i] = vi;
    }
  }
}

}  // namespace

extern "C" {

int svit_grad_sqnorm(const svit_optim_tensor* table, const int32_t* chunk_tensor, const int64_t* chunk_start, int nchunks,
                     int chunk, float* out, void* stream) {
  if (!table || !chunk_tensor || !chunk_start || !out || nchunks < 0 || chunk < 1) return SVIT_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  SVIT_CUDA(cudaMemsetAsync(out, 0, sizeof(float), st));
  if (nchunks == 0) return 0;
  const int grid = nchunks < svit_num_sms() * 8 ? nchunks : svit_num_sms() * 8;
  grad_sqnorm_kernel<<<grid, 256, 0, st>>>(table, chunk_tensor, chunk_start, nchunks, chunk, out);
  SVIT_CHECK_LAUNCH();
  return 0;
}

int svit_adamw_step(const svit_optim_tensor* table, const int32_t* chunk_tensor, const int64_t* chunk_start, int nchunks,
                    int chunk, float lr, float beta1, float beta2, float eps, int step, float max_norm, const float* sqnorm,
                    void* stream) {
  if (!table || !chunk_tensor || !chunk_start || nchunks < 0 || chunk < 1 || step < 1) return SVIT_EINVAL;
  if (nchunks == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float sqrt_bc2 = sqrtf(1.f - powf(beta2, (float)step));
  const int grid = nchunks < svit_num_sms() * 8 ? nchunks : svit_num_sms() * 8;
  adamw_kernel<<<grid, 256, 0, st>>>(table, chunk_tensor, chunk_start, nchunks, chunk, lr, beta1, beta2, eps, bc1, sqrt_bc2,
                                     max_norm, sqnorm, nullptr);
  SVIT_CHECK_LAUNCH();
  return 0;
}

int svit_adamw_step_dev(const svit_optim_tensor* table, const int32_t* chunk_tensor, const int64_t* chunk_start, int nchunks,
                        int chunk, const float* hyper, float beta1, float beta2, float eps, float max_norm,
                        const float* sqnorm, void* stream) {
  if (!table || !chunk_tensor || !chunk_start || !hyper || nchunks < 0 || chunk < 1) return SVIT_EINVAL;
  if (nchunks == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = nchunks < svit_num_sms() * 8 ? nchunks : svit_num_sms() * 8;
  adamw_kernel<<<grid, 256, 0, st>>>(table, chunk_tensor, chunk_start, nchunks, chunk, 0.f, beta1, beta2, eps, 1.f, 1.f,
                                     max_norm, sqnorm, hyper);
  SVIT_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
