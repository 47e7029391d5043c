// One-launch Adam over all parameters of the learner (pipeline.py:232-257 uses torch.optim.Adam; gomoku/run_training.py:110).
// torch's fused multi-tensor Adam walks the 181 parameter tensors of the Gomoku net in 64 K-element chunks -- 113 CTAs
// spread over five launches, 0.2 ms of a 6.5 ms training step; this kernel takes the same tensors (parameter, gradient
// and the optimizer's own exp_avg / exp_avg_sq state, in place) through a chunk table: one launch, every SM busy.
// Same arithmetic as torch.optim.Adam (L2 weight decay added to the gradient, bias-corrected step size, eps outside the
// square root), float32.
#include "common.cuh"
#include <cmath>

namespace mz {
namespace {

constexpr int kAdamChunk = 4096;
constexpr int kAdamThreads = 256;

__global__ void __launch_bounds__(kAdamThreads) adam_kernel(const mz_adam_tensor* __restrict__ tensors, const int32_t* __restrict__ chunk_tensor,
                                                            const int64_t* __restrict__ chunk_start, const float* __restrict__ step_dev,
                                                            const float* __restrict__ lr_dev, double beta1d, double beta2d, float eps,
                                                            float weight_decay) {
  // 1 - beta is formed in float32 from the rounded beta, as torch's fused kernel does (1 - float(0.999) is off by 5e-5 of
  // the 0.001 that weighs every new squared gradient: forming it in double would be "more exact" and 1e-5 away from torch)
  const float beta1 = (float)beta1d, beta2 = (float)beta2d, omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
  __shared__ float s_step_size, s_bc2_sqrt;
  if (threadIdx.x == 0) {
    const double step = (double)*step_dev;                          // already incremented by the caller
    const double bc1 = 1.0 - pow(beta1d, step), bc2 = 1.0 - pow(beta2d, step);
    s_step_size = (float)((double)*lr_dev / bc1);
    s_bc2_sqrt = (float)sqrt(bc2);
  }
  const mz_adam_tensor t = tensors[chunk_tensor[blockIdx.x]];
  const int64_t begin = chunk_start[blockIdx.x];
  const int64_t end = begin + kAdamChunk < t.n ? begin + kAdamChunk : t.n;
  __syncthreads();
  const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
#pragma unroll 4
  for (int64_t i = begin + threadIdx.x; i < end; i += kAdamThreads) {
    float p = t.p[i];
    float g = t.g[i];
    if (weight_decay != 0.0f) g = fmaf(weight_decay, p, g);
    float m = t.m[i], v = t.v[i];
    m = m + omb1 * (g - m);                                         // lerp(exp_avg, grad, 1 - beta1)
    v = beta2 * v + omb2 * g * g;
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    p -= step_size * (m / denom);
    t.p[i] = p; t.m[i] = m; t.v[i] = v;
  }
}

}  // namespace
}  // namespace mz

extern "C" {

int mz_adam_chunk_elements(void) { return mz::kAdamChunk; }

int mz_adam_step(const mz_adam_tensor* tensors_dev, const int32_t* chunk_tensor_dev, const int64_t* chunk_start_dev, int32_t n_chunks,
                 const float* step_dev, const float* lr_dev, double beta1, double beta2, double eps, double weight_decay,
                 mz_stream stream) {
  MZ_CHECK_ARG(tensors_dev != nullptr && chunk_tensor_dev != nullptr && chunk_start_dev != nullptr && step_dev != nullptr && lr_dev != nullptr,
               "mz_adam_step: NULL argument");
  MZ_CHECK_ARG(n_chunks > 0, "mz_adam_step: no chunks");
  mz::adam_kernel<<<n_chunks, mz::kAdamThreads, 0, static_cast<cudaStream_t>(stream)>>>(tensors_dev, chunk_tensor_dev, chunk_start_dev, step_dev,
                                                                                      lr_dev, beta1, beta2, (float)eps, (float)weight_decay);
  MZ_LAUNCH_CHECK("adam_kernel");
  return MZ_OK;
}

}  // extern "C"
