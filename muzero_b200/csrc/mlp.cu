// Fused fp32 inference kernels for the MuZeroMLPNet family (network.py:140-267).
//
// One CTA runs the whole initial_inference / recurrent_inference chain for a
// tile of 32 rows (trees): hidden-state gather by slot index, one-hot action as
// a weight-row gather, both Linear-ReLU-Linear stacks, min-max normalisation
// (util.py:31-36), softmax and the support->scalar transform (util.py:70-93)
// never leave shared memory.  These nets are 0.18-0.40 MFLOP per row: the path
// is launch/latency bound, so the design goal is ONE launch per network call.
#include "common.cuh"
#include "net.cuh"

namespace mz {

constexpr int kRows = 32;
constexpr int kThreads = 256;

struct MlpDev {
  int in_dim, A, P, HD, Sv, Sr;
  int in_pad;   // in_dim rounded up to 4
  // transposed [in, out] weights
  const float *rep1_wt, *rep1_b, *rep2_wt, *rep2_b;
  const float *dyn1_wt, *dyn1_b, *dyn2_wt, *dyn2_b;
  const float *rew1_wt, *rew1_b, *rew2_wt, *rew2_w, *rew2_b;   // *_w: original [out, in] layout for tiny heads
  const float *pol1_wt, *pol1_b, *pol2_wt, *pol2_w, *pol2_b;
  const float *val1_wt, *val1_b, *val2_wt, *val2_w, *val2_b;
};

// out[r][j] = act(bias[j] + sum_k in[r][k] * Wt[k][j] (+ Wt[extra_row[r]][j]))
// thread -> (column j, row group g); RPT rows per thread held in registers.
template <int RPT>
__device__ void dense_cols(const float* __restrict__ Wt, const float* __restrict__ bias, const float* in, int ldin,
                           int K, int OUT, bool relu, float* out, int ldout, const int* extra_row) {
  constexpr int G = kRows / RPT;
  const int items = OUT * G;
  for (int idx = threadIdx.x; idx < items; idx += kThreads) {
    const int j = idx % OUT, g = idx / OUT;
    const float* x0 = in + (size_t)g * RPT * ldin;
    float acc[RPT];
    const float b = bias[j];
#pragma unroll
    for (int r = 0; r < RPT; ++r) acc[r] = b;
    int k = 0;
    const int K4 = K & ~3;
#pragma unroll 2
    for (; k < K4; k += 4) {
      const float w0 = __ldg(Wt + (size_t)(k + 0) * OUT + j);
      const float w1 = __ldg(Wt + (size_t)(k + 1) * OUT + j);
      const float w2 = __ldg(Wt + (size_t)(k + 2) * OUT + j);
      const float w3 = __ldg(Wt + (size_t)(k + 3) * OUT + j);
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const float4 xv = *reinterpret_cast<const float4*>(x0 + (size_t)r * ldin + k);
        acc[r] = fmaf(xv.x, w0, acc[r]);
        acc[r] = fmaf(xv.y, w1, acc[r]);
        acc[r] = fmaf(xv.z, w2, acc[r]);
        acc[r] = fmaf(xv.w, w3, acc[r]);
      }
    }
    for (; k < K; ++k) {
      const float w = __ldg(Wt + (size_t)k * OUT + j);
#pragma unroll
      for (int r = 0; r < RPT; ++r) acc[r] = fmaf(x0[(size_t)r * ldin + k], w, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      float v = acc[r];
      if (extra_row) v += __ldg(Wt + (size_t)extra_row[g * RPT + r] * OUT + j);
      if (relu) v = fmaxf(v, 0.0f);
      out[(size_t)(g * RPT + r) * ldout + j] = v;
    }
  }
}

// tiny heads (OUT < 16): a warp per (row, output), lanes split K; W in [OUT, K] layout
__device__ void dense_tiny(const float* __restrict__ W, const float* __restrict__ bias, const float* in, int ldin,
                           int K, int OUT, float* out, int ldout) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int item = warp; item < kRows * OUT; item += kThreads / 32) {
    const int r = item / OUT, j = item % OUT;
    float acc = 0.0f;
    for (int k = lane; k < K; k += 32) acc = fmaf(in[(size_t)r * ldin + k], __ldg(W + (size_t)j * K + k), acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[(size_t)r * ldout + j] = acc + bias[j];
  }
}

__device__ void dense(const float* Wt, const float* W, const float* bias, const float* in, int ldin, int K, int OUT,
                      bool relu, float* out, int ldout, const int* extra_row = nullptr) {
  if (OUT < 16 && W != nullptr && !relu && extra_row == nullptr) dense_tiny(W, bias, in, ldin, K, OUT, out, ldout);
  else if (OUT >= 256) dense_cols<32>(Wt, bias, in, ldin, K, OUT, relu, out, ldout, extra_row);
  else if (OUT >= 128) dense_cols<16>(Wt, bias, in, ldin, K, OUT, relu, out, ldout, extra_row);
  else if (OUT >= 64) dense_cols<8>(Wt, bias, in, ldin, K, OUT, relu, out, ldout, extra_row);
  else if (OUT >= 32) dense_cols<4>(Wt, bias, in, ldin, K, OUT, relu, out, ldout, extra_row);
  else dense_cols<2>(Wt, bias, in, ldin, K, OUT, relu, out, ldout, extra_row);
  __syncthreads();
}

// util.py:25-28 signed_parabolic, float32, same operation order as the torch expression
__device__ __forceinline__ float signed_parabolic(float x) {
  const float eps = 1e-3f;
  float z = __fadd_rn(1.0f, __fmul_rn(4.0f * eps, __fadd_rn(eps + 1.0f, fabsf(x))));
  z = __fsqrt_rn(z);
  z = __fdiv_rn(__fdiv_rn(z, 2.0f), eps);
  z = __fsub_rn(z, 500.0f);                           // 1 / 2 / eps
  const float r = __fsub_rn(__fmul_rn(z, z), 1.0f);
  return x > 0.0f ? r : (x < 0.0f ? -r : 0.0f * r);
}

// util.py:70-93: softmax over the support, expectation, inverse value transform.  One warp per row.
__device__ void support_to_scalar(const float* logits, int ld, int S, float* out_global, int row0, int batch) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < kRows; r += kThreads / 32) {
    if (row0 + r >= batch) continue;
    const float* l = logits + (size_t)r * ld;
    if (S == 1) {                                     // scalar head, network.py:126-134
      if (lane == 0) out_global[row0 + r] = l[0];
      continue;
    }
    float m = -INFINITY;
    for (int i = lane; i < S; i += 32) m = fmaxf(m, l[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const int maxv = (S - 1) / 2;
    const float step = S > 1 ? (float)(2 * maxv) / (float)(S - 1) : 0.0f;
    float den = 0.0f, num = 0.0f;
    for (int i = lane; i < S; i += 32) {
      const float e = expf(l[i] - m);
      den += e;
      num += e * ((float)(-maxv) + step * (float)i);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      den += __shfl_xor_sync(0xffffffffu, den, o);
      num += __shfl_xor_sync(0xffffffffu, num, o);
    }
    if (lane == 0) out_global[row0 + r] = signed_parabolic(num / den);
  }
  __syncthreads();
}

__device__ void softmax_rows(const float* logits, int ld, int A, float* out_global, int row0, int batch) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < kRows; r += kThreads / 32) {
    if (row0 + r >= batch) continue;
    const float* l = logits + (size_t)r * ld;
    float m = -INFINITY;
    for (int i = lane; i < A; i += 32) m = fmaxf(m, l[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float den = 0.0f;
    for (int i = lane; i < A; i += 32) den += expf(l[i] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
    for (int i = lane; i < A; i += 32) out_global[(size_t)(row0 + r) * A + i] = expf(l[i] - m) / den;
  }
  __syncthreads();
}

// util.py:31-36 over the feature dimension; writes the normalised state to smem and to its slot
__device__ void normalise_and_store(const float* hraw, float* hn, int HD, float* hidden_out,
                                    const int32_t* dst_index, int row0, int batch) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < kRows; r += kThreads / 32) {
    const float* h = hraw + (size_t)r * HD;
    float mn = INFINITY, mx = -INFINITY;
    for (int i = lane; i < HD; i += 32) { mn = fminf(mn, h[i]); mx = fmaxf(mx, h[i]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    const float den = __fadd_rn(__fsub_rn(mx, mn), 1e-8f);
    const bool live = row0 + r < batch;
    float* dst = nullptr;
    if (live) dst = hidden_out + (size_t)(dst_index ? dst_index[row0 + r] : row0 + r) * HD;
    for (int i = lane; i < HD; i += 32) {
      const float v = __fdiv_rn(__fsub_rn(h[i], mn), den);
      hn[(size_t)r * HD + i] = v;
      if (live) dst[i] = v;
    }
  }
  __syncthreads();
}

struct Smem {
  float *x, *h1, *hraw, *hn, *lg;
  int *arow;
  int ldx, ldl;
};
__device__ Smem carve(const MlpDev& n, float* base) {
  Smem s;
  s.ldx = max(n.in_pad, n.HD);
  s.ldl = (max(max(n.A, n.Sv), n.Sr) + 3) & ~3;
  s.x = base;
  s.h1 = s.x + kRows * s.ldx;
  s.hraw = s.h1 + kRows * n.P;
  s.hn = s.hraw + kRows * n.HD;
  s.lg = s.hn + kRows * n.HD;
  s.arow = reinterpret_cast<int*>(s.lg + kRows * s.ldl);
  return s;
}
size_t smem_bytes(const MlpDev& n) {
  const int ldx = n.in_pad > n.HD ? n.in_pad : n.HD;
  int m = n.A > n.Sv ? n.A : n.Sv;
  m = m > n.Sr ? m : n.Sr;
  const int ldl = (m + 3) & ~3;
  return sizeof(float) * ((size_t)kRows * ldx + (size_t)kRows * n.P + 2 * (size_t)kRows * n.HD + (size_t)kRows * ldl) +
         sizeof(int) * kRows;
}

extern __shared__ __align__(16) float mlp_smem[];

// prediction heads on the normalised state in s.hn (network.py:223-233)
__device__ void prediction_heads(const MlpDev& n, const Smem& s, float* pi_probs, float* value, int row0, int batch) {
  if (pi_probs != nullptr) {
    dense(n.pol1_wt, nullptr, n.pol1_b, s.hn, n.HD, n.HD, n.P, true, s.h1, n.P);
    dense(n.pol2_wt, n.pol2_w, n.pol2_b, s.h1, n.P, n.P, n.A, false, s.lg, s.ldl);
    softmax_rows(s.lg, s.ldl, n.A, pi_probs, row0, batch);
  }
  dense(n.val1_wt, nullptr, n.val1_b, s.hn, n.HD, n.HD, n.P, true, s.h1, n.P);
  dense(n.val2_wt, n.val2_w, n.val2_b, s.h1, n.P, n.P, n.Sv, false, s.lg, s.ldl);
  support_to_scalar(s.lg, s.ldl, n.Sv, value, row0, batch);
}

__global__ void __launch_bounds__(kThreads)
mlp_initial_kernel(MlpDev n, int batch, const float* __restrict__ obs, float* __restrict__ hidden_out,
                   const int32_t* __restrict__ dst_index, float* __restrict__ pi_probs, float* __restrict__ value) {
  const Smem s = carve(n, mlp_smem);
  const int row0 = blockIdx.x * kRows;
  for (int i = threadIdx.x; i < kRows * s.ldx; i += kThreads) {
    const int r = i / s.ldx, k = i % s.ldx;
    s.x[i] = (row0 + r < batch && k < n.in_dim) ? obs[(size_t)(row0 + r) * n.in_dim + k] : 0.0f;
  }
  __syncthreads();
  dense(n.rep1_wt, nullptr, n.rep1_b, s.x, s.ldx, n.in_dim, n.P, true, s.h1, n.P);
  dense(n.rep2_wt, nullptr, n.rep2_b, s.h1, n.P, n.P, n.HD, false, s.hraw, n.HD);
  normalise_and_store(s.hraw, s.hn, n.HD, hidden_out, dst_index, row0, batch);
  prediction_heads(n, s, pi_probs, value, row0, batch);
}

__global__ void __launch_bounds__(kThreads)
mlp_recurrent_kernel(MlpDev n, int batch, const float* __restrict__ hidden_in, const int32_t* __restrict__ src_index,
                     const int32_t* __restrict__ action, float* __restrict__ hidden_out,
                     const int32_t* __restrict__ dst_index, float* __restrict__ reward, float* __restrict__ value,
                     float* __restrict__ pi_probs) {
  const Smem s = carve(n, mlp_smem);
  const int row0 = blockIdx.x * kRows;
  // leaf gather: parent hidden state by slot index + the action as a weight-row index
  for (int i = threadIdx.x; i < kRows * (n.HD / 4); i += kThreads) {
    const int r = i / (n.HD / 4), k4 = i % (n.HD / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < batch) {
      const size_t slot = src_index ? src_index[row0 + r] : row0 + r;
      v = *reinterpret_cast<const float4*>(hidden_in + slot * n.HD + 4 * k4);
    }
    *reinterpret_cast<float4*>(s.x + (size_t)r * s.ldx + 4 * k4) = v;
  }
  if (threadIdx.x < kRows) {
    const int r = threadIdx.x;
    int a = (row0 + r < batch) ? action[row0 + r] : 0;
    a = min(max(a, 0), n.A - 1);
    s.arow[r] = n.HD + a;      // one-hot(action) concatenated after the state (network.py:191-193)
  }
  __syncthreads();
  dense(n.dyn1_wt, nullptr, n.dyn1_b, s.x, s.ldx, n.HD, n.P, true, s.h1, n.P, s.arow);
  dense(n.dyn2_wt, nullptr, n.dyn2_b, s.h1, n.P, n.P, n.HD, false, s.hraw, n.HD);
  // reward head reads the UN-normalised state (network.py:195-196)
  dense(n.rew1_wt, nullptr, n.rew1_b, s.hraw, n.HD, n.HD, n.P, true, s.h1, n.P);
  dense(n.rew2_wt, n.rew2_w, n.rew2_b, s.h1, n.P, n.P, n.Sr, false, s.lg, s.ldl);
  support_to_scalar(s.lg, s.ldl, n.Sr, reward, row0, batch);
  normalise_and_store(s.hraw, s.hn, n.HD, hidden_out, dst_index, row0, batch);
  prediction_heads(n, s, pi_probs, value, row0, batch);
}

__global__ void transpose_kernel(const float* __restrict__ w, float* __restrict__ wt, int out, int in) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < out * in) {
    const int o = i / in, k = i % in;
    wt[(size_t)k * out + o] = w[i];
  }
}

// ---------------------------------------------------------------------------
struct MlpNet : NetImpl {
  MlpDev d;
  size_t smem;

  int initial(int batch, const float* obs, void* hidden_out, const int32_t* dst_index, float* pi_probs,
              float* value, cudaStream_t st) override {
    prof_mark(kProfMlp, st);
    mlp_initial_kernel<<<(batch + kRows - 1) / kRows, kThreads, smem, st>>>(d, batch, obs, (float*)hidden_out,
                                                                           dst_index, pi_probs, value);
    prof_mark(-1, st);
    MZ_LAUNCH_CHECK("mlp_initial_kernel");
    return MZ_OK;
  }
  int recurrent(int batch, const void* hidden_in, const int32_t* src_index, const int32_t* action, void* hidden_out,
                const int32_t* dst_index, float* reward, float* value, float* pi_probs, cudaStream_t st) override {
    prof_mark(kProfMlp, st);
    mlp_recurrent_kernel<<<(batch + kRows - 1) / kRows, kThreads, smem, st>>>(
        d, batch, (const float*)hidden_in, src_index, action, (float*)hidden_out, dst_index, reward, value, pi_probs);
    prof_mark(-1, st);
    MZ_LAUNCH_CHECK("mlp_recurrent_kernel");
    return MZ_OK;
  }
};

static int mlp_dims(const mz_net_config& c, int* in_dim) {
  MZ_CHECK_ARG(c.hidden_dim > 0 && c.hidden_dim % 4 == 0, "hidden_dim must be a positive multiple of 4, got %d",
               c.hidden_dim);
  MZ_CHECK_ARG(c.num_planes > 0 && c.num_actions > 0 && c.value_support > 0 && c.reward_support > 0,
               "bad MLP dimensions");
  *in_dim = c.in_channels * c.in_h * c.in_w;
  MZ_CHECK_ARG(*in_dim > 0, "bad observation shape");
  return MZ_OK;
}

int mlp_hidden_bytes(const mz_net_config& c, int32_t* bytes) {
  int in_dim;
  int rc = mlp_dims(c, &in_dim);
  if (rc) return rc;
  *bytes = c.hidden_dim * 4;
  return MZ_OK;
}

// shapes of the 20 state_dict tensors, in order (network.py:143-147,170-180,210-220)
static void mlp_shapes(const mz_net_config& c, int in_dim, int (*sh)[2]) {
  const int P = c.num_planes, H = c.hidden_dim, A = c.num_actions;
  const int s[20][2] = {{P, in_dim}, {P, 1}, {H, P}, {H, 1},
                        {P, H + A}, {P, 1}, {H, P}, {H, 1},
                        {P, H}, {P, 1}, {c.reward_support, P}, {c.reward_support, 1},
                        {P, H}, {P, 1}, {A, P}, {A, 1},
                        {P, H}, {P, 1}, {c.value_support, P}, {c.value_support, 1}};
  memcpy(sh, s, sizeof(s));
}

int mlp_arena_bytes(const mz_net_config& c, size_t* bytes) {
  int in_dim;
  int rc = mlp_dims(c, &in_dim);
  if (rc) return rc;
  int sh[20][2];
  mlp_shapes(c, in_dim, sh);
  size_t tot = 0;
  for (int i = 0; i < 20; ++i) tot += 2 * align_up((size_t)sh[i][0] * sh[i][1] * 4, 256);   // [out,in] + transposed
  *bytes = tot;
  return MZ_OK;
}

int mlp_create(const mz_net_config& c, const float* const* w, int nw, void* arena, size_t arena_bytes, NetImpl** out) {
  int in_dim;
  int rc = mlp_dims(c, &in_dim);
  if (rc) return rc;
  MZ_CHECK_ARG(nw == 20, "MuZeroMLPNet has 20 state_dict tensors, got %d", nw);
  size_t need;
  mlp_arena_bytes(c, &need);
  if (arena_bytes < need) { set_error("net arena too small: %zu < %zu", arena_bytes, need); return MZ_ENOMEM; }
  int sh[20][2];
  mlp_shapes(c, in_dim, sh);
  char* p = (char*)arena;
  const float* orig[20];
  const float* tr[20];
  for (int i = 0; i < 20; ++i) {
    const size_t n = (size_t)sh[i][0] * sh[i][1];
    float* o = (float*)p; p += align_up(n * 4, 256);
    float* t = (float*)p; p += align_up(n * 4, 256);
    MZ_CUDA(cudaMemcpy(o, w[i], n * 4, cudaMemcpyDeviceToDevice));
    if (sh[i][1] > 1) {
      transpose_kernel<<<(int)((n + 255) / 256), 256>>>(o, t, sh[i][0], sh[i][1]);
      MZ_LAUNCH_CHECK("transpose_kernel");
    }
    orig[i] = o; tr[i] = t;
  }
  MZ_CUDA(cudaDeviceSynchronize());
  MlpNet* net = new MlpNet();
  MlpDev& d = net->d;
  d.in_dim = in_dim; d.A = c.num_actions; d.P = c.num_planes; d.HD = c.hidden_dim;
  d.Sv = c.value_support; d.Sr = c.reward_support; d.in_pad = (in_dim + 3) & ~3;
  d.rep1_wt = tr[0];  d.rep1_b = orig[1];  d.rep2_wt = tr[2];  d.rep2_b = orig[3];
  d.dyn1_wt = tr[4];  d.dyn1_b = orig[5];  d.dyn2_wt = tr[6];  d.dyn2_b = orig[7];
  d.rew1_wt = tr[8];  d.rew1_b = orig[9];  d.rew2_wt = tr[10]; d.rew2_w = orig[10]; d.rew2_b = orig[11];
  d.pol1_wt = tr[12]; d.pol1_b = orig[13]; d.pol2_wt = tr[14]; d.pol2_w = orig[14]; d.pol2_b = orig[15];
  d.val1_wt = tr[16]; d.val1_b = orig[17]; d.val2_wt = tr[18]; d.val2_w = orig[18]; d.val2_b = orig[19];
  net->smem = smem_bytes(d);
  if (net->smem > 227 * 1024) {
    set_error("MLP too wide for the fused kernel: needs %zu bytes of shared memory", net->smem);
    delete net;
    return MZ_EINVAL;
  }
  MZ_CUDA(cudaFuncSetAttribute(mlp_initial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)net->smem));
  MZ_CUDA(cudaFuncSetAttribute(mlp_recurrent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)net->smem));
  *out = net;
  return MZ_OK;
}

}  // namespace mz
