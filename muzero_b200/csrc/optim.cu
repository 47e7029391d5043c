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

// ---------------------------------------------------------------------------------------------------------------
// The 1x1 convolutions of the heads (network.py:398-470: Conv2d(planes, 1 or 2, kernel_size=1)) over the K calls'
// stacked tower outputs: x [N][C][HW] float32 (N = K * B boards), w [M][C], M <= 4.  cuDNN's route for such an output goes
// through NCHW <-> NHWC transposes of the whole 26 MB input and a split-K weight gradient (~75 us per head and step);
// these three kernels each stream the tensor once (~8 us).
// ---------------------------------------------------------------------------------------------------------------
namespace mz {
namespace {

constexpr int kHeadMaxM = 4;
constexpr int kHeadMaxC = 256;

__global__ void __launch_bounds__(256) head_conv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y,
                                                            long long NP, int C, int HW, int M) {
  __shared__ float s_w[kHeadMaxM * kHeadMaxC];
  for (int i = threadIdx.x; i < M * C; i += blockDim.x) s_w[i] = w[i];
  __syncthreads();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // (board n, position)
  if (t >= NP) return;
  const long long n = t / HW;
  const int pos = (int)(t - n * HW);
  const float* xp = x + n * (long long)C * HW + pos;
  float acc[kHeadMaxM] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 8
  for (int c = 0; c < C; ++c) {
    const float v = __ldg(xp + (long long)c * HW);
#pragma unroll
    for (int m = 0; m < kHeadMaxM; ++m)
      if (m < M) acc[m] = fmaf(s_w[m * C + c], v, acc[m]);
  }
  for (int m = 0; m < M; ++m) y[(n * M + m) * HW + pos] = acc[m];
}

// dx[n][c][pos] = sum_m w[m][c] dy[n][m][pos]
__global__ void __launch_bounds__(256) head_conv_dx_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx,
                                                           long long NP, int C, int HW, int M) {
  __shared__ float s_w[kHeadMaxM * kHeadMaxC];
  for (int i = threadIdx.x; i < M * C; i += blockDim.x) s_w[i] = w[i];
  __syncthreads();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= NP) return;
  const long long n = t / HW;
  const int pos = (int)(t - n * HW);
  float g[kHeadMaxM] = {0.0f, 0.0f, 0.0f, 0.0f};
  for (int m = 0; m < M; ++m) g[m] = dy[(n * M + m) * HW + pos];
  float* dp = dx + n * (long long)C * HW + pos;
#pragma unroll 8
  for (int c = 0; c < C; ++c) {
    float v = 0.0f;
#pragma unroll
    for (int m = 0; m < kHeadMaxM; ++m)
      if (m < M) v = fmaf(s_w[m * C + c], g[m], v);
    dp[(long long)c * HW] = v;
  }
}

// dw[m][c] += sum over this block's boards and all positions of dy[n][m][pos] x[n][c][pos]; thread = channel c.
// Per-block partials go to `partial` [blocks][M][C]; head_conv_dw_reduce_kernel adds them in block order (no float atomics:
// the result does not depend on scheduling).
__global__ void __launch_bounds__(kHeadMaxC) head_conv_dw_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                 float* __restrict__ partial, int N, int C, int HW, int M,
                                                                 int boards_per_block) {
  extern __shared__ float s_dy[];                                   // [M][HW] of the current board
  const int c = threadIdx.x;
  float acc[kHeadMaxM] = {0.0f, 0.0f, 0.0f, 0.0f};
  const int n0 = blockIdx.x * boards_per_block;
  for (int n = n0; n < n0 + boards_per_block && n < N; ++n) {
    __syncthreads();
    for (int i = threadIdx.x; i < M * HW; i += blockDim.x) s_dy[i] = dy[(long long)n * M * HW + i];
    __syncthreads();
    if (c < C) {
      const float* xp = x + ((long long)n * C + c) * HW;
      for (int pos = 0; pos < HW; ++pos) {
        const float v = __ldg(xp + pos);
#pragma unroll
        for (int m = 0; m < kHeadMaxM; ++m)
          if (m < M) acc[m] = fmaf(s_dy[m * HW + pos], v, acc[m]);
      }
    }
  }
  if (c < C)
    for (int m = 0; m < M; ++m) partial[((long long)blockIdx.x * M + m) * C + c] = acc[m];
}

__global__ void head_conv_dw_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int blocks, int MC) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= MC) return;
  float s = 0.0f;
  for (int b = 0; b < blocks; ++b) s += partial[(long long)b * MC + i];
  dw[i] = s;
}

}  // namespace
}  // namespace mz

extern "C" {

int mz_head_conv_forward(const float* x, const float* w, float* y, int64_t n, int32_t c, int32_t hw, int32_t m, mz_stream stream) {
  MZ_CHECK_ARG(x != nullptr && w != nullptr && y != nullptr, "mz_head_conv_forward: NULL argument");
  MZ_CHECK_ARG(n > 0 && c > 0 && c <= mz::kHeadMaxC && hw > 0 && m > 0 && m <= mz::kHeadMaxM, "mz_head_conv_forward: shape n=%lld c=%d hw=%d m=%d not supported",
               (long long)n, c, hw, m);
  const long long np = (long long)n * hw;
  mz::head_conv_fwd_kernel<<<(unsigned)((np + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, w, y, np, c, hw, m);
  MZ_LAUNCH_CHECK("head_conv_fwd_kernel");
  return MZ_OK;
}

size_t mz_head_conv_scratch_bytes(int64_t n, int32_t c, int32_t m) {
  const int64_t blocks = (n + 3) / 4;
  return (size_t)blocks * (size_t)m * (size_t)c * sizeof(float);
}

int mz_head_conv_backward(const float* x, const float* w, const float* dy, float* dx, float* dw, void* scratch, int64_t n, int32_t c, int32_t hw,
                          int32_t m, mz_stream stream) {
  MZ_CHECK_ARG(x != nullptr && w != nullptr && dy != nullptr && dx != nullptr && dw != nullptr && scratch != nullptr,
               "mz_head_conv_backward: NULL argument");
  MZ_CHECK_ARG(n > 0 && c > 0 && c <= mz::kHeadMaxC && hw > 0 && m > 0 && m <= mz::kHeadMaxM && (size_t)m * hw * 4 <= 40 * 1024,
               "mz_head_conv_backward: shape n=%lld c=%d hw=%d m=%d not supported", (long long)n, c, hw, m);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long np = (long long)n * hw;
  mz::head_conv_dx_kernel<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(dy, w, dx, np, c, hw, m);
  MZ_LAUNCH_CHECK("head_conv_dx_kernel");
  const int bpb = 4, blocks = (int)((n + bpb - 1) / bpb);
  mz::head_conv_dw_kernel<<<blocks, mz::kHeadMaxC, (size_t)m * hw * sizeof(float), st>>>(x, dy, static_cast<float*>(scratch), (int)n, c, hw, m, bpb);
  MZ_LAUNCH_CHECK("head_conv_dw_kernel");
  mz::head_conv_dw_reduce_kernel<<<(m * c + 127) / 128, 128, 0, st>>>(static_cast<const float*>(scratch), dw, blocks, m * c);
  MZ_LAUNCH_CHECK("head_conv_dw_reduce_kernel");
  return MZ_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// The rest of a head (network.py:398-470: BatchNorm2d(mid) in train mode, ReLU, Flatten, Linear(mid * hw, O)) over the K
// calls' stacked 1x1-convolution outputs y [calls * B][mid][hw], every call with its OWN batch statistics -- what
// network.head_over_calls computes with ~33 small PyTorch kernels per head and step, as six.  mid <= 4, mid * hw <= 1024.
// Statistics and the Linear weight gradient are summed in fixed orders (no float atomics): results do not depend on
// scheduling.
// ---------------------------------------------------------------------------------------------------------------
namespace mz {
namespace {

constexpr int kHeadMaxJ = 1024;

constexpr int kStatThreads = 1024;
__device__ __forceinline__ double block_sum_double(double v, double* red) {        // kStatThreads threads, fixed order
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < kStatThreads / 32; ++w) s += red[w];
  return s;
}

// block (m, call): mean and 1 / sqrt(var + eps) of channel m over the call's B * hw values; saved [calls][mid][3] also keeps
// the biased variance for the running statistics
__global__ void __launch_bounds__(kStatThreads) head_bn_stats_kernel(const float* __restrict__ y, float* __restrict__ saved, int B, int mid,
                                                                     int HW, float eps) {
  __shared__ double red[kStatThreads / 32];
  const int m = blockIdx.x, call = blockIdx.y;
  const long long n0 = (long long)call * B;
  const int cnt = B * HW;
  double s = 0.0, q = 0.0;                          // one pass, double accumulators: E[x^2] - mean^2 loses nothing here
  for (int i = threadIdx.x; i < cnt; i += kStatThreads) {
    const int b = i / HW, pos = i - b * HW;
    const double v = (double)y[((n0 + b) * mid + m) * HW + pos];
    s += v;
    q += v * v;
  }
  const double mean = block_sum_double(s, red) / cnt;
  double var = block_sum_double(q, red) / cnt - mean * mean;
  var = var > 0.0 ? var : 0.0;
  if (threadIdx.x == 0) {
    float* o = saved + ((size_t)call * mid + m) * 3;
    o[0] = (float)mean; o[1] = (float)(1.0 / sqrt(var + (double)eps)); o[2] = (float)var;
  }
}

// block = board n: z = relu(bn(y)) [J = mid * hw] -> out = W z + b [O]; block 0 also moves the running statistics through
// the calls' updates in call order
__global__ void __launch_bounds__(256) head_tail_fwd_kernel(const float* __restrict__ y, const float* __restrict__ saved,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const float* __restrict__ W, const float* __restrict__ bias, float* __restrict__ z,
                                                            float* __restrict__ out, float* __restrict__ rmean, float* __restrict__ rvar, int B,
                                                            int calls, int mid, int HW, int O, float momentum) {
  __shared__ float s_z[kHeadMaxJ];
  const long long n = blockIdx.x;
  const int call = (int)(n / B), J = mid * HW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (n == 0 && threadIdx.x < mid) {
    const int m = threadIdx.x;
    const double cnt = (double)B * HW;
    float rm = rmean[m], rv = rvar[m];
    for (int c = 0; c < calls; ++c) {
      const float* sv = saved + ((size_t)c * mid + m) * 3;
      rm = (1.0f - momentum) * rm + momentum * sv[0];
      rv = (1.0f - momentum) * rv + momentum * (float)((double)sv[2] * (cnt / (cnt - 1.0)));
    }
    rmean[m] = rm; rvar[m] = rv;
  }
  for (int j = threadIdx.x; j < J; j += 256) {
    const int m = j / HW;
    const float* sv = saved + ((size_t)call * mid + m) * 3;
    const float v = fmaxf((y[n * J + j] - sv[0]) * sv[1] * gamma[m] + beta[m], 0.0f);
    s_z[j] = v;
    z[n * J + j] = v;
  }
  __syncthreads();
  for (int o = warp; o < O; o += 8) {
    const float* wr = W + (size_t)o * J;
    float acc = 0.0f;
    for (int j = lane; j < J; j += 32) acc = fmaf(__ldg(wr + j), s_z[j], acc);
    for (int k = 16; k > 0; k >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, k);
    if (lane == 0) out[n * O + o] = acc + bias[o];
  }
}

// block = board n: dz = W^T dout, masked by z > 0 -> dzr [J]
__global__ void __launch_bounds__(256) head_tail_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ W, const float* __restrict__ z,
                                                            float* __restrict__ dzr, int J, int O) {
  __shared__ float s_d[128];
  const long long n = blockIdx.x;
  for (int o = threadIdx.x; o < O; o += 256) s_d[o] = dout[n * O + o];
  __syncthreads();
  for (int j = threadIdx.x; j < J; j += 256) {
    float acc = 0.0f;
    for (int o = 0; o < O; ++o) acc = fmaf(__ldg(W + (size_t)o * J + j), s_d[o], acc);
    dzr[n * J + j] = z[n * J + j] > 0.0f ? acc : 0.0f;
  }
}

// block (m, call): S1 = sum dzr, S2 = sum dzr * xhat over the call's B * hw values -> sums [calls][mid][2]
__global__ void __launch_bounds__(kStatThreads) head_bn_bwd_stats_kernel(const float* __restrict__ dzr, const float* __restrict__ y,
                                                                         const float* __restrict__ saved, float* __restrict__ sums, int B,
                                                                         int mid, int HW) {
  __shared__ double red[kStatThreads / 32];
  const int m = blockIdx.x, call = blockIdx.y;
  const long long n0 = (long long)call * B;
  const int cnt = B * HW;
  const float* sv = saved + ((size_t)call * mid + m) * 3;
  const float mean = sv[0], is = sv[1];
  double s1 = 0.0, s2 = 0.0;
  for (int i = threadIdx.x; i < cnt; i += kStatThreads) {
    const int b = i / HW, pos = i - b * HW;
    const size_t at = ((n0 + b) * mid + m) * HW + pos;
    const float d = dzr[at];
    s1 += (double)d;
    s2 += (double)(d * ((y[at] - mean) * is));
  }
  const double t1 = block_sum_double(s1, red);
  const double t2 = block_sum_double(s2, red);
  if (threadIdx.x == 0) { sums[((size_t)call * mid + m) * 2] = (float)t1; sums[((size_t)call * mid + m) * 2 + 1] = (float)t2; }
}

// dy = gamma invstd (dzr - S1 / cnt - xhat S2 / cnt); block 0: dgamma[m] = sum_calls S2, dbeta[m] = sum_calls S1 (call order)
__global__ void __launch_bounds__(256) head_bn_bwd_apply_kernel(const float* __restrict__ dzr, const float* __restrict__ y,
                                                                const float* __restrict__ saved, const float* __restrict__ sums,
                                                                const float* __restrict__ gamma, float* __restrict__ dy, float* __restrict__ dgamma,
                                                                float* __restrict__ dbeta, long long total, int B, int calls, int mid, int HW) {
  if (blockIdx.x == 0 && threadIdx.x < mid) {
    float g = 0.0f, b = 0.0f;
    for (int c = calls - 1; c >= 0; --c) {            // the order separate calls' backward passes would run in
      b += sums[((size_t)c * mid + threadIdx.x) * 2];
      g += sums[((size_t)c * mid + threadIdx.x) * 2 + 1];
    }
    dgamma[threadIdx.x] = g; dbeta[threadIdx.x] = b;
  }
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const long long n = i / ((long long)mid * HW);
  const int m = (int)((i / HW) % mid), call = (int)(n / B);
  const float* sv = saved + ((size_t)call * mid + m) * 3;
  const float* sm = sums + ((size_t)call * mid + m) * 2;
  const float inv_cnt = 1.0f / ((float)B * HW);
  const float xh = (y[i] - sv[0]) * sv[1];
  dy[i] = gamma[m] * sv[1] * (dzr[i] - sm[0] * inv_cnt - xh * sm[1] * inv_cnt);
}

}  // namespace
}  // namespace mz

extern "C" {

int mz_head_tail_forward(const float* y, const float* gamma, const float* beta, const float* w, const float* bias, float* running_mean,
                         float* running_var, float* saved, float* z, float* out, int32_t calls, int32_t b, int32_t mid, int32_t hw, int32_t o,
                         float eps, float momentum, mz_stream stream) {
  MZ_CHECK_ARG(y && gamma && beta && w && bias && running_mean && running_var && saved && z && out, "mz_head_tail_forward: NULL argument");
  MZ_CHECK_ARG(calls > 0 && b > 1 && mid > 0 && mid <= mz::kHeadMaxM && hw > 0 && mid * hw <= mz::kHeadMaxJ && o > 0 && o <= 128,
               "mz_head_tail_forward: shape calls=%d b=%d mid=%d hw=%d o=%d not supported", calls, b, mid, hw, o);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  mz::head_bn_stats_kernel<<<dim3(mid, calls), mz::kStatThreads, 0, st>>>(y, saved, b, mid, hw, eps);
  MZ_LAUNCH_CHECK("head_bn_stats_kernel");
  mz::head_tail_fwd_kernel<<<calls * b, 256, 0, st>>>(y, saved, gamma, beta, w, bias, z, out, running_mean, running_var, b, calls, mid, hw, o, momentum);
  MZ_LAUNCH_CHECK("head_tail_fwd_kernel");
  return MZ_OK;
}

int mz_head_tail_backward(const float* dout, const float* y, const float* gamma, const float* w, const float* saved, const float* z, float* dzr,
                          float* sums, float* dy, float* dgamma, float* dbeta, int32_t calls, int32_t b, int32_t mid, int32_t hw, int32_t o,
                          mz_stream stream) {
  MZ_CHECK_ARG(dout && y && gamma && w && saved && z && dzr && sums && dy && dgamma && dbeta, "mz_head_tail_backward: NULL argument");
  MZ_CHECK_ARG(calls > 0 && b > 1 && mid > 0 && mid <= mz::kHeadMaxM && hw > 0 && mid * hw <= mz::kHeadMaxJ && o > 0 && o <= 128,
               "mz_head_tail_backward: shape calls=%d b=%d mid=%d hw=%d o=%d not supported", calls, b, mid, hw, o);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n = (long long)calls * b;
  const int J = mid * hw;
  mz::head_tail_bwd_kernel<<<(unsigned)n, 256, 0, st>>>(dout, w, z, dzr, J, o);
  MZ_LAUNCH_CHECK("head_tail_bwd_kernel");
  mz::head_bn_bwd_stats_kernel<<<dim3(mid, calls), mz::kStatThreads, 0, st>>>(dzr, y, saved, sums, b, mid, hw);
  MZ_LAUNCH_CHECK("head_bn_bwd_stats_kernel");
  const long long total = n * J;
  mz::head_bn_bwd_apply_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dzr, y, saved, sums, gamma, dy, dgamma, dbeta, total, b, calls, mid, hw);
  MZ_LAUNCH_CHECK("head_bn_bwd_apply_kernel");
  return MZ_OK;
}

}  // extern "C"
