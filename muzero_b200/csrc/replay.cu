// Device-resident replay: ring-buffer scatter, batch gather and the reference's sampling
// (muzero/replay.py:84-105) on the GPU.  Bit-exactness contract, same as the tree kernels: the index stream is the
// one numpy's legacy MT19937 produces -- RandomState.uniform for the uniform path (replay.py:90), RandomState.choice
// with p for the prioritized path (replay.py:96: float64 running sum of the float32 probabilities, division by the
// last entry, searchsorted side='right') -- and float32 / float64 operations are issued one IEEE instruction at a time.
#include <vector>

#include "common.cuh"
#include "rng.cuh"

namespace mz {
namespace {

constexpr unsigned kAll = 0xffffffffu;

// np.sum of a contiguous float32 array (numpy pairwise_sum), one thread.  8 independent accumulators per 128-block
// keep the adds pipelined; the recursion of numpy (halves rounded down to a multiple of 8) is replayed with an
// explicit stack.
__device__ float np_pairwise_sum_f32(const float* __restrict__ a, long long n) {
  struct Frame { long long off, n; int state; float left; };
  Frame st[48];
  int sp = 0;
  st[0] = {0, n, 0, 0.0f};
  float ret = 0.0f;
  while (sp >= 0) {
    Frame& f = st[sp];
    if (f.state == 0) {
      if (f.n < 8) {
        float res = -0.0f;
        for (long long i = 0; i < f.n; ++i) res = __fadd_rn(res, a[f.off + i]);
        ret = res; --sp;
      } else if (f.n <= 128) {
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = a[f.off + j];
        long long i = 8;
        const long long lim = f.n - (f.n % 8);
        for (; i < lim; i += 8) {
#pragma unroll
          for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[f.off + i + j]);
        }
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                              __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (; i < f.n; ++i) res = __fadd_rn(res, a[f.off + i]);
        ret = res; --sp;
      } else {
        long long n2 = f.n / 2;
        n2 -= n2 % 8;
        f.state = 1;
        st[sp + 1] = {f.off, n2, 0, 0.0f};
        ++sp;
      }
    } else if (f.state == 1) {
      f.left = ret;
      long long n2 = f.n / 2;
      n2 -= n2 % 8;
      f.state = 2;
      st[sp + 1] = {f.off + n2, f.n - n2, 0, 0.0f};
      ++sp;
    } else {
      ret = __fadd_rn(f.left, ret);
      --sp;
    }
  }
  return ret;
}

// x ** e the way numpy computes float32 ** python float for the exponents with an exact fast path (1, 2, 0.5);
// other exponents go through powf (CUDA and the host libm agree to ~2 ulp)
__device__ __forceinline__ float np_pow_f32(float x, float e) {
  if (e == 1.0f) return x;
  if (e == 2.0f) return __fmul_rn(x, x);
  if (e == 0.5f) return __fsqrt_rn(x);
  return powf(x, e);
}

// ---- uniform: idx = trunc(size * u), u consecutive doubles of the replay's own stream ---------------------------
__global__ void __launch_bounds__(32) sample_uniform_kernel(long long size, int batch, uint32_t* key, int* pos,
                                                             long long* idx, float* w) {
  const int lane = threadIdx.x;
  WarpRng rng;
  rng.load(key, pos, lane);
  for (int k = 0; k < batch; ++k) {
    const double u = rng.next_double();
    const double x = __dadd_rn(0.0, __dmul_rn((double)size, u));
    if (lane == (k & 31)) { idx[k] = (long long)x; w[k] = 1.0f; }
  }
  __syncwarp();
  rng.store(pos);
}

// ---- prioritized, step 1: probs = p ** alpha / sum (float32), cdf = running float64 sum / last -------------------
__global__ void pow_kernel(const float* __restrict__ prio, float* __restrict__ pw, long long n, float alpha) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    pw[i] = np_pow_f32(prio[i], alpha);
}
// The float32 pairwise total, bit-exact and parallel: numpy's recursion (halves rounded down to a multiple of 8, leaves
// of at most 128 elements) is a FIXED tree, so its sub-trees can be summed independently.  Thread `path` (kTotalDepth
// bits, first split = most significant bit) walks the recursion to its node and sums that node's segment with the
// sequential routine above (which is numpy's recursion on the segment); one thread then combines the partial sums along
// the top of the same tree.  A node that numpy would not split any further (<= 128 elements) before kTotalDepth splits
// belongs to the path whose remaining bits are zero.
constexpr int kTotalDepth = 10;
__device__ __forceinline__ long long np_split(long long n) { long long n2 = n / 2; return n2 - n2 % 8; }
__global__ void total_leaves_kernel(const float* __restrict__ pw, long long n, float* __restrict__ partial) {
  const int path = blockIdx.x * blockDim.x + threadIdx.x;
  if (path >= (1 << kTotalDepth)) return;
  long long off = 0, len = n;
  for (int level = 0; level < kTotalDepth; ++level) {
    if (len <= 128) {                                     // numpy stops here: owned by the all-zero continuation
      if (path & ((1 << (kTotalDepth - level)) - 1)) return;
      break;
    }
    const long long n2 = np_split(len);
    if ((path >> (kTotalDepth - 1 - level)) & 1) { off += n2; len -= n2; } else { len = n2; }
  }
  partial[path] = np_pairwise_sum_f32(pw + off, len);
}
__global__ void total_combine_kernel(const float* __restrict__ partial, long long n, float* total) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  // depth-first walk of the top kTotalDepth levels with an explicit stack (state: 0 enter, 1 left done, 2 right done)
  struct Frame { long long len; int path, level, state; float left; };
  Frame st[kTotalDepth + 2];
  int sp = 0;
  st[0] = {n, 0, 0, 0, 0.0f};
  float ret = 0.0f;
  while (sp >= 0) {
    Frame& f = st[sp];
    if (f.state == 0) {
      if (f.level == kTotalDepth || f.len <= 128) { ret = partial[f.path << (kTotalDepth - f.level)]; --sp; continue; }
      f.state = 1;
      st[sp + 1] = {np_split(f.len), f.path << 1, f.level + 1, 0, 0.0f};
      ++sp;
    } else if (f.state == 1) {
      f.left = ret;
      f.state = 2;
      st[sp + 1] = {f.len - np_split(f.len), (f.path << 1) | 1, f.level + 1, 0, 0.0f};
      ++sp;
    } else {
      ret = __fadd_rn(f.left, ret);
      --sp;
    }
  }
  *total = ret;
}
__global__ void probs_kernel(float* __restrict__ pw, long long n, const float* total) {
  const float s = *total;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    pw[i] = __fdiv_rn(pw[i], s);
}
// The float64 running sum cdf[i] = cdf[i-1] + p[i] rounds after every addition, which orders it -- unless no addition
// rounds at all.  Every p[i] is a float32 (24-bit significand) and every partial sum stays below 2, so if each nonzero
// p[i] is at least 2^-29 all its bits are multiples of 2^-52 = ulp(1.x) and EVERY partial sum is exactly representable:
// the additions are exact, hence associative, and a parallel scan returns the very doubles the sequential loop does.
// exact_check_kernel raises a flag when some p[i] is smaller than that (or the total could reach 2); the sequential
// kernel then runs instead.  (A replay of 10^6 items has p ~ 10^-6 = 2^-20.)
constexpr int kScanChunk = 2048;       // elements per block
__global__ void exact_check_kernel(const float* __restrict__ probs, long long n, int* flag) {
  bool bad = false;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float p = probs[i];
    bad |= !(p == 0.0f || (p >= 1.862645149230957e-09f && p < 1.0f));      // 2^-29 <= p < 1, or exactly 0
  }
  if (__any_sync(kAll, bad) && (threadIdx.x & 31) == 0) atomicExch(flag, 1);
}
__global__ void __launch_bounds__(256) scan_block_totals_kernel(const float* __restrict__ probs, long long n,
                                                                double* __restrict__ block_total, const int* flag) {
  if (*flag) return;
  __shared__ double sh[8];
  const long long base = (long long)blockIdx.x * kScanChunk;
  double acc = 0.0;
  for (int j = threadIdx.x; j < kScanChunk; j += 256) {
    const long long i = base + j;
    if (i < n) acc += (double)probs[i];                   // exact additions: any order
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kAll, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    block_total[blockIdx.x] = t;
  }
}
__global__ void scan_offsets_kernel(double* __restrict__ block_total, int nblocks, const int* flag) {
  if (*flag || threadIdx.x != 0 || blockIdx.x != 0) return;
  double acc = 0.0;
  for (int b = 0; b < nblocks; ++b) { const double t = block_total[b]; block_total[b] = acc; acc += t; }   // exclusive
  // the total must stay below 2 for the exactness argument (it is ~1)
}
__global__ void __launch_bounds__(256) scan_write_kernel(const float* __restrict__ probs, long long n,
                                                         const double* __restrict__ block_offset,
                                                         double* __restrict__ cdf, const int* flag) {
  if (*flag) return;
  __shared__ double sh[256];
  constexpr int PER = kScanChunk / 256;                    // consecutive elements per thread
  const long long first = (long long)blockIdx.x * kScanChunk + (long long)threadIdx.x * PER;
  double v[PER];
  double acc = 0.0;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const long long i = first + j;
    acc += (i < n) ? (double)probs[i] : 0.0;
    v[j] = acc;
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  // exclusive prefix of the per-thread totals (256 entries: a simple Hillis-Steele scan; every addition is exact)
  for (int o = 1; o < 256; o <<= 1) {
    const double add = threadIdx.x >= o ? sh[threadIdx.x - o] : 0.0;
    __syncthreads();
    sh[threadIdx.x] += add;
    __syncthreads();
  }
  const double before = block_offset[blockIdx.x] + (threadIdx.x ? sh[threadIdx.x - 1] : 0.0);
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const long long i = first + j;
    if (i < n) cdf[i] = before + v[j];
  }
}
// sequential form (the general case: some addition may round)
__global__ void cumsum_kernel(const float* __restrict__ probs, long long n, double* __restrict__ cdf, const int* flag) {
  if (threadIdx.x != 0 || blockIdx.x != 0 || *flag == 0) return;
  double acc = 0.0;
  long long i = 0;
  for (; i + 4 <= n; i += 4) {          // loads run ahead of the dependent adds
    const float a = probs[i], b = probs[i + 1], c = probs[i + 2], d = probs[i + 3];
    acc = __dadd_rn(acc, (double)a); cdf[i] = acc;
    acc = __dadd_rn(acc, (double)b); cdf[i + 1] = acc;
    acc = __dadd_rn(acc, (double)c); cdf[i + 2] = acc;
    acc = __dadd_rn(acc, (double)d); cdf[i + 3] = acc;
  }
  for (; i < n; ++i) { acc = __dadd_rn(acc, (double)probs[i]); cdf[i] = acc; }
}
__global__ void cdf_norm_kernel(double* __restrict__ cdf, long long n) {
  const double last = cdf[n - 1];
  // every entry but the last first (they all divide by the ORIGINAL last entry); the last one becomes exactly 1
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n - 1; i += (long long)gridDim.x * blockDim.x)
    cdf[i] = __ddiv_rn(cdf[i], last);
}
__global__ void cdf_last_kernel(double* __restrict__ cdf, long long n) { cdf[n - 1] = __ddiv_rn(cdf[n - 1], cdf[n - 1]); }

// ---- prioritized, step 2: draws on the global stream, searchsorted 'right', importance weights ------------------
__global__ void __launch_bounds__(32) sample_cdf_kernel(long long size, int batch, const double* __restrict__ cdf,
                                                         const float* __restrict__ probs, float beta, uint32_t* key,
                                                         int* pos, long long* idx, float* w) {
  const int lane = threadIdx.x;
  WarpRng rng;
  rng.load(key, pos, lane);
  // (1.0 / size) is a python float; dividing it by a float32 array makes it a float32 scalar first (NEP 50)
  const float uni = (float)__ddiv_rn(1.0, (double)size);
  float wmax = -INFINITY;
  for (int k0 = 0; k0 < batch; k0 += 32) {
    double u = 0.0;
    for (int j = 0; j < 32 && k0 + j < batch; ++j) {       // the stream is sequential: every lane draws every double
      const double d = rng.next_double();
      if (j == lane) u = d;
    }
    const int k = k0 + lane;
    if (k < batch) {
      long long lo = 0, hi = size;                          // first index with cdf[i] > u
      while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
      }
      idx[k] = lo;
      const float wk = np_pow_f32(__fdiv_rn(uni, probs[lo < size ? lo : size - 1]), beta);
      w[k] = wk;
      wmax = fmaxf(wmax, wk);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(kAll, wmax, o));
  __syncwarp();
  for (int k = lane; k < batch; k += 32) w[k] = __fdiv_rn(w[k], wmax);
  rng.store(pos);
}

// ---- rows in, rows out -------------------------------------------------------------------------------------------
// dst[(start + i) % capacity] = src[i]   (replay.py:76-79, n items at once)
__global__ void scatter_rows_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, long long n,
                                    long long row_bytes, long long start, long long capacity) {
  const long long units = row_bytes % 16 == 0 ? row_bytes / 16 : row_bytes;
  const bool vec = row_bytes % 16 == 0;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n * units;
       g += (long long)gridDim.x * blockDim.x) {
    const long long i = g / units, c = g % units;
    const long long slot = (start + i) % capacity;
    if (vec) reinterpret_cast<int4*>(dst + slot * row_bytes)[c] = reinterpret_cast<const int4*>(src + i * row_bytes)[c];
    else dst[slot * row_bytes + c] = src[i * row_bytes + c];
  }
}
// dst[i] = src[idx[i]]   (replay.py:81-83 + the np.stack of replay.py:103)
__global__ void gather_rows_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                   const long long* __restrict__ idx, long long n, long long row_bytes) {
  const long long units = row_bytes % 16 == 0 ? row_bytes / 16 : row_bytes;
  const bool vec = row_bytes % 16 == 0;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n * units;
       g += (long long)gridDim.x * blockDim.x) {
    const long long i = g / units, c = g % units;
    const long long slot = idx[i];
    if (vec) reinterpret_cast<int4*>(dst + i * row_bytes)[c] = reinterpret_cast<const int4*>(src + slot * row_bytes)[c];
    else dst[i * row_bytes + c] = src[slot * row_bytes + c];
  }
}
// priorities[idx[i]] = p[i], in order (later duplicates win, like the reference's loop, replay.py:113-114)
__global__ void update_priorities_kernel(float* __restrict__ prio, const long long* __restrict__ idx,
                                         const float* __restrict__ p, int n) {
  if (threadIdx.x == 0 && blockIdx.x == 0)
    for (int i = 0; i < n; ++i) prio[idx[i]] = p[i];
}

inline int blocks_for(long long work) {
  long long b = (work + 255) / 256;
  return (int)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b));
}

}  // namespace
}  // namespace mz

using namespace mz;

extern "C" int mz_replay_sample_uniform(int64_t size, int32_t batch, uint32_t* rng_key, int32_t* rng_pos,
                                        int64_t* out_index, float* out_weight, mz_stream stream) {
  MZ_CHECK_ARG(rng_key && rng_pos && out_index && out_weight, "NULL argument");
  MZ_CHECK_ARG(size > 0 && batch > 0, "size and batch must be positive");
  sample_uniform_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((long long)size, batch, rng_key, rng_pos,
                                                            (long long*)out_index, out_weight);
  MZ_LAUNCH_CHECK("sample_uniform_kernel");
  return MZ_OK;
}

extern "C" int mz_replay_sample_prioritized(int64_t size, int32_t batch, const float* priorities,
                                            float priority_exponent, float importance_exponent, uint32_t* rng_key,
                                            int32_t* rng_pos, float* scratch_probs, double* scratch_cdf,
                                            int64_t* out_index, float* out_weight, mz_stream stream) {
  MZ_CHECK_ARG(priorities && rng_key && rng_pos && scratch_probs && scratch_cdf && out_index && out_weight,
               "NULL argument");
  MZ_CHECK_ARG(size > 0 && batch > 0, "size and batch must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)size;
  float* total = scratch_probs + n;                 // one float past the probabilities
  // caller-owned scratch (the library allocates nothing): behind the n probabilities and the total come the partial
  // sums of the parallel pairwise total and the exactness flag; behind the n cdf entries the block offsets of the scan
  float* partial = total + 1;
  int* flag = reinterpret_cast<int*>(partial + (1 << kTotalDepth));
  double* block_total = scratch_cdf + n;
  const int nblocks = (int)((n + kScanChunk - 1) / kScanChunk);
  pow_kernel<<<blocks_for(n), 256, 0, st>>>(priorities, scratch_probs, n, priority_exponent);
  MZ_LAUNCH_CHECK("pow_kernel");
  total_leaves_kernel<<<(1 << kTotalDepth) / 128, 128, 0, st>>>(scratch_probs, n, partial);
  MZ_LAUNCH_CHECK("total_leaves_kernel");
  total_combine_kernel<<<1, 32, 0, st>>>(partial, n, total);
  MZ_LAUNCH_CHECK("total_combine_kernel");
  probs_kernel<<<blocks_for(n), 256, 0, st>>>(scratch_probs, n, total);
  MZ_LAUNCH_CHECK("probs_kernel");
  MZ_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
  exact_check_kernel<<<blocks_for(n), 256, 0, st>>>(scratch_probs, n, flag);
  MZ_LAUNCH_CHECK("exact_check_kernel");
  scan_block_totals_kernel<<<nblocks, 256, 0, st>>>(scratch_probs, n, block_total, flag);
  MZ_LAUNCH_CHECK("scan_block_totals_kernel");
  scan_offsets_kernel<<<1, 32, 0, st>>>(block_total, nblocks, flag);
  MZ_LAUNCH_CHECK("scan_offsets_kernel");
  scan_write_kernel<<<nblocks, 256, 0, st>>>(scratch_probs, n, block_total, scratch_cdf, flag);
  MZ_LAUNCH_CHECK("scan_write_kernel");
  cumsum_kernel<<<1, 32, 0, st>>>(scratch_probs, n, scratch_cdf, flag);
  MZ_LAUNCH_CHECK("cumsum_kernel");
  cdf_norm_kernel<<<blocks_for(n), 256, 0, st>>>(scratch_cdf, n);
  MZ_LAUNCH_CHECK("cdf_norm_kernel");
  cdf_last_kernel<<<1, 1, 0, st>>>(scratch_cdf, n);
  MZ_LAUNCH_CHECK("cdf_last_kernel");
  sample_cdf_kernel<<<1, 32, 0, st>>>(n, batch, scratch_cdf, scratch_probs, importance_exponent, rng_key, rng_pos,
                                      (long long*)out_index, out_weight);
  MZ_LAUNCH_CHECK("sample_cdf_kernel");
  return MZ_OK;
}

extern "C" int mz_replay_scatter(const void* rows, void* storage, int64_t n, int64_t row_bytes, int64_t start,
                                 int64_t capacity, mz_stream stream) {
  MZ_CHECK_ARG(rows && storage, "NULL argument");
  MZ_CHECK_ARG(n >= 0 && row_bytes > 0 && capacity > 0 && start >= 0, "bad sizes");
  MZ_CHECK_ARG(n <= capacity, "more rows (%lld) than capacity (%lld) in one call", (long long)n, (long long)capacity);
  if (n == 0) return MZ_OK;
  const long long units = row_bytes % 16 == 0 ? row_bytes / 16 : row_bytes;
  scatter_rows_kernel<<<blocks_for(n * units), 256, 0, (cudaStream_t)stream>>>(
      (const uint8_t*)rows, (uint8_t*)storage, n, row_bytes, start, capacity);
  MZ_LAUNCH_CHECK("scatter_rows_kernel");
  return MZ_OK;
}

extern "C" int mz_replay_gather(const void* storage, const int64_t* index, void* rows, int64_t n, int64_t row_bytes,
                                mz_stream stream) {
  MZ_CHECK_ARG(storage && index && rows, "NULL argument");
  MZ_CHECK_ARG(n > 0 && row_bytes > 0, "bad sizes");
  const long long units = row_bytes % 16 == 0 ? row_bytes / 16 : row_bytes;
  gather_rows_kernel<<<blocks_for(n * units), 256, 0, (cudaStream_t)stream>>>(
      (const uint8_t*)storage, (uint8_t*)rows, (const long long*)index, n, row_bytes);
  MZ_LAUNCH_CHECK("gather_rows_kernel");
  return MZ_OK;
}

extern "C" int mz_replay_update_priorities(float* priorities, const int64_t* index, const float* values, int32_t n,
                                           mz_stream stream) {
  MZ_CHECK_ARG(priorities && index && values, "NULL argument");
  MZ_CHECK_ARG(n > 0, "n must be positive");
  update_priorities_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(priorities, (const long long*)index, values, n);
  MZ_LAUNCH_CHECK("update_priorities_kernel");
  return MZ_OK;
}
