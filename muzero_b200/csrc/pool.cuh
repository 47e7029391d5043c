// Device-side view of a search pool and the per-tree ROOT preparation (Dirichlet draw, noise mix, masking,
// renormalisation, root expansion, MinMaxStats reset), shared by the tree kernels (mcts.cu) and by the prediction
// epilogues of the initial inference (mlp.cu, conv.cu), which run it fused behind the policy softmax
// (BASELINE.json north_star: "Dirichlet root noise fused into the prediction epilogue").
#pragma once
#include "common.cuh"
#include "rng.cuh"

namespace mz {

struct PoolDev {
  int B, A, S, max_nodes;
  int board;
  double discount, dp;
  HotEdge* hot;      // [B][max_nodes][A] {child << 16 | N, child_Q}: all the descent reads
  double* ew;        // f64 [B][max_nodes][A] Node.W        (backup only; defined where N > 0)
  float* er;         // f32 [B][max_nodes][A] Node.reward   (backup only; defined where N > 0)
  double* prior;
  double* rootW;
  int* rootN;
  double* minmax;
  int* count;
  int *leaf_parent, *leaf_action, *leaf_depth, *src_slot, *dst_slot;
  uint32_t* path;
  int *node_parent, *node_move;
  float* node_value;
  uint32_t* rng_key;
  int* rng_pos;
  float *reward, *value;
  int* error;
  unsigned long long* stats;
  const double* T;
  uint8_t* same_player;
  double* root_reward;
  uint8_t* f32_prior;
  double bound_min, bound_max;
  int has_bounds;
  unsigned* work;   // {next tree, finished CTAs} of the confined tree kernel
  int timing;   // MZ_TREE_TIMING: stats[4..6] = min / max block start and max block end of the tree kernels (%globaltimer)
};

PoolDev pool_dev(const mz_pool* h);       // mcts.cu

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ HotEdge hot_empty() { return make_uint2((uint32_t)kNoChild << 16, 0u); }

// ---------------------------------------------------------------------------
// numpy pairwise summation (np.sum of a contiguous 1-D array), sequential
// ---------------------------------------------------------------------------
template <typename T> struct Add;
template <> struct Add<float>  { static __device__ float  f(float a, float b)   { return __fadd_rn(a, b); } };
template <> struct Add<double> { static __device__ double f(double a, double b) { return __dadd_rn(a, b); } };

// `a` holds doubles; on the float32 path they are exactly float32 values and T = float.
template <typename T>
__device__ T pairwise_sum(const double* a, int n) {
  if (n < 8) {
    T res = (T)(-0.0);
    for (int i = 0; i < n; ++i) res = Add<T>::f(res, (T)a[i]);
    return res;
  }
  if (n <= 128) {
    T r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = (T)a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = Add<T>::f(r[j], (T)a[i + j]);
    }
    T res = Add<T>::f(Add<T>::f(Add<T>::f(r[0], r[1]), Add<T>::f(r[2], r[3])),
                      Add<T>::f(Add<T>::f(r[4], r[5]), Add<T>::f(r[6], r[7])));
    for (; i < n; ++i) res = Add<T>::f(res, (T)a[i]);
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  return Add<T>::f(pairwise_sum<T>(a, n2), pairwise_sum<T>(a + n2, n - n2));
}

// ---------------------------------------------------------------------------
// numpy legacy_standard_gamma / dirichlet on a tree's MT19937 stream (mcts.py:244-245)
// ---------------------------------------------------------------------------
template <typename Rng>
__device__ __noinline__ double legacy_gamma(Rng& rng, double shape) {
  if (shape == 1.0) return -log(__dsub_rn(1.0, rng.next_double()));
  if (shape == 0.0) return 0.0;
  while (true) {
    const double u = rng.next_double();
    const double v = -log(__dsub_rn(1.0, rng.next_double()));
    if (u <= __dsub_rn(1.0, shape)) {
      const double xx = pow(u, __ddiv_rn(1.0, shape));
      if (xx <= v) return xx;
    } else {
      const double y = -log(__ddiv_rn(__dsub_rn(1.0, u), shape));
      const double xx = pow(__dadd_rn(__dsub_rn(1.0, shape), __dmul_rn(shape, y)), __ddiv_rn(1.0, shape));
      if (xx <= __dadd_rn(v, y)) return xx;
    }
  }
}

// Dirichlet(alpha, ..., alpha) for tree t, written to o[0..A): every lane runs the identical sequential sampler.
__device__ __forceinline__ void dirichlet_tree(const PoolDev& p, int t, int lane, double alpha, double* o) {
  WarpRng rng;
  rng.load(p.rng_key + (size_t)t * 624, p.rng_pos + t, lane);
  double acc = 0.0;
  for (int a = 0; a < p.A; ++a) {
    const double g = legacy_gamma(rng, alpha);
    acc = __dadd_rn(acc, g);
    if (lane == 0) o[a] = g;
  }
  __syncwarp();
  const double inv = __ddiv_rn(1.0, acc);
  for (int a = lane; a < p.A; a += 32) o[a] = __dmul_rn(o[a], inv);
  rng.store(p.rng_pos + t);
  __syncwarp();
}

// Root preparation of tree t by one warp (replaces mcts.py:353-367 with 244-246 and 283-299): prior =
// (1 - eps) * pi + eps * noise (float32 product + float64 product, see reset_kernel's history), masked and
// renormalised with numpy's pairwise sum; root row zeroed, MinMaxStats reset.  pi / noise / mask point at THIS tree's
// rows; noise == nullptr keeps the prior float32.
__device__ __forceinline__ void root_setup_tree(const PoolDev& p, int t, int lane, const float* pi, const double* noise,
                                                double eps, float one_minus_eps_f32, const uint8_t* mask,
                                                const int32_t* players_t, const float* root_reward_t) {
  const int A = p.A;
  double* P = p.prior + (size_t)t * A;
  const bool f32p = (noise == nullptr);
  for (int a = lane; a < A; a += 32) {
    const float pf = pi[a];
    double pd;
    if (f32p) {
      pd = (double)pf;
    } else {
      // (1 - eps) * prob is a float32 product (weak python scalar); eps * noise is float64
      pd = __dadd_rn((double)__fmul_rn(one_minus_eps_f32, pf), __dmul_rn(eps, noise[a]));
    }
    if (mask != nullptr && mask[a] == 0) pd = 0.0;
    P[a] = pd;
  }
  __syncwarp();
  if (mask != nullptr) {
    // sequential on every lane (identical results), cheaper than a broadcast for A <= a few hundred
    if (f32p) {
      const float s = pairwise_sum<float>(P, A);
      __syncwarp();
      if (s > 0.0f)
        for (int a = lane; a < A; a += 32) P[a] = (double)__fdiv_rn((float)P[a], s);
    } else {
      const double s = pairwise_sum<double>(P, A);
      __syncwarp();
      if (s > 0.0)
        for (int a = lane; a < A; a += 32) P[a] = __ddiv_rn(P[a], s);
    }
  }
  // root expansion: row 0 zeroed, no children yet
  HotEdge* row = p.hot + (size_t)t * p.max_nodes * A;
  for (int a = lane; a < A; a += 32) row[a] = hot_empty();
  if (lane == 0) {
    p.rootW[t] = 0.0;
    p.rootN[t] = 0;
    p.minmax[2 * t] = p.has_bounds ? p.bound_min : __longlong_as_double(0x7ff0000000000000LL);
    p.minmax[2 * t + 1] = p.has_bounds ? p.bound_max : __longlong_as_double(0xfff0000000000000LL);
    p.count[t] = 1;
    p.node_parent[(size_t)t * p.max_nodes] = -1;
    p.node_move[(size_t)t * p.max_nodes] = -1;
    p.same_player[t] = (players_t == nullptr) ? 1 : (players_t[0] == players_t[1]);
    p.root_reward[t] = (root_reward_t == nullptr) ? 0.0 : (double)root_reward_t[0];
    p.f32_prior[t] = f32p ? 1 : 0;
    p.leaf_depth[t] = 0;
  }
}

// What a prediction epilogue needs to prepare the roots of the trees whose policy it just computed
// (mz_net_initial_search): row i of the inference is tree i of `pool`.
struct RootSetup {
  PoolDev pool;
  int enabled;                 // 0: plain initial inference
  int noise_mode;              // 0: no noise (prior stays float32), 1: `noise` given, 2: drawn here from the trees' streams
  double* noise;               // f64 [B, A]: read (mode 1) or written (mode 2)
  double alpha, eps;
  float one_minus_eps_f32;
  const uint8_t* mask;         // u8 [B, A] or nullptr
  const int32_t* players;      // i32 [B, 2] or nullptr
};

// one warp, tree t, pi = this tree's softmax row (global memory, written by this warp)
__device__ __forceinline__ void root_setup_fused(const RootSetup& rs, int t, int lane, const float* pi_row) {
  const PoolDev& p = rs.pool;
  if (t >= p.B) return;
  __syncwarp();
  double* nz = rs.noise_mode ? rs.noise + (size_t)t * p.A : nullptr;
  if (rs.noise_mode == 2) dirichlet_tree(p, t, lane, rs.alpha, nz);
  root_setup_tree(p, t, lane, pi_row, nz, rs.eps, rs.one_minus_eps_f32,
                  rs.mask ? rs.mask + (size_t)t * p.A : nullptr, rs.players ? rs.players + 2 * t : nullptr, nullptr);
}

}  // namespace mz
