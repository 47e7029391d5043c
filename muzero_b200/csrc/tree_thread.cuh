// Thread-per-tree forms of the tree operations (select / expand + backup) and the scalar arithmetic they share with
// the warp-per-tree kernels of mcts.cu.  Used by tree_thread_kernel (tiny action spaces) and by the persistent
// per-search kernel of the MLP networks (mlp.cu), where the thread that owns row i of a network tile also owns tree i.
// Same arithmetic, instruction for instruction, as the warp kernels (child_Q cache, div_by_count, child_q, the backup
// recurrence), same MT19937 stream (sequential twist), same memory -- only the work split differs.
#pragma once
#include "pool.cuh"

namespace mz {

// order-preserving map float -> uint32 (so the warp maximum is one redux.sync instruction)
__device__ __forceinline__ uint32_t f2ord(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Node.child_Q of one edge (mcts.py:159-178), one IEEE instruction per reference operation: evaluated by the
// BACKUP (off the descent's dependent chain, lanes in parallel) and cached as float32 per edge.  The value depends on
// the edge's own (W, N, reward) and on the tree's min-max bounds, so the backup refreshes the edges of its path, or
// every visited edge of the tree when it moved a bound.  Why: while the conv tower of the other sub-batch runs
// tcgen05.mma on the same SM, FP64 throughput collapses (a dependent DFMA chain is 9.5x slower, measured by
// tools/coresident_probe.py), so the descent keeps as little float64 arithmetic as bit-exactness allows.
__device__ __forceinline__ float child_q(double W, float R, uint32_t N, double dp, bool norm, double lo, double range) {
  if (N == 0) return 0.0f;
  double v = __dadd_rn((double)R, __dmul_rn(dp, __ddiv_rn(W, (double)N)));
  if (norm) v = __ddiv_rn(__dsub_rn(v, lo), range);
  return __double2float_rn(v);
}

// Correctly rounded float64 division without the ~35-instruction IEEE division sequence: with y = RN(1/b),
//   q0 = RN(a*y);  r0 = a - b*q0 (one FMA, exact);  q1 = RN(q0 + r0*y)
// q1 == RN(a/b) for every integer divisor b < 2^16 (a quotient by a small integer is never closer than 2^-17 ulp to a
// rounding boundary, the error of q0 + r0*y before rounding is ~2^-53 ulp).  Operands outside a +-2^400 exponent
// window (and zeros, infinities, NaNs) take the IEEE path.  tools/fastdiv_check.c searches 2e8 operand pairs over all
// 65535 divisors for a mismatch (none; tests/test_abi.py runs a 5 % sample).  Used for y = T[N] / (n + 1) on the
// descent's dependent chain; everything else that divides (child_Q) runs in the backup with IEEE divisions.
// out of line on purpose: inlined, ptxas if-converts the IEEE sequence back into the hot path
static __device__ __noinline__ double ieee_div_cold(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ bool fastdiv_window(double a) {
  const uint32_t e = ((uint32_t)__double2hiint(a) >> 20) & 0x7ffu;
  return e - 623u <= 800u;
}
__device__ __forceinline__ double div_by_count(double a, int b, double y) {
  const double db = (double)b;
  const double q0 = __dmul_rn(a, y);
  const double r0 = __fma_rn(-db, q0, a);
  const double q1 = __fma_rn(r0, y, q0);
  if (fastdiv_window(a)) return q1;
  return ieee_div_cold(a, db);
}
struct TreeThreadStats {
  unsigned depth = 0, draws = 0, twists = 0;
};

// TA = compile-time bound on the number of actions (the per-action arrays live in registers).  Leaves the leaf in the
// pool's LEAF_* / SRC_SLOT / DST_SLOT / PATH views like the warp kernels do and returns (leaf parent node, action).
template <int TA>
__device__ __forceinline__ int2 select_tree_thread(const PoolDev& p, const int t, const double* sT, const double* sR,
                                                   TreeThreadStats& st) {
  const int A = p.A;
  const HotEdge* tree = p.hot + (size_t)t * p.max_nodes * A;
  const double* __restrict__ P = p.prior + (size_t)t * A;
  const bool f32p = p.f32_prior[t] != 0;
  uint32_t* pth = p.path + (size_t)t * p.max_nodes;
  ThreadRng rng;
  rng.load(p.rng_key + (size_t)t * 624, p.rng_pos + t);
  double pr[TA];
  float prf[TA];
#pragma unroll
  for (int a = 0; a < TA; ++a) { pr[a] = P[a < A ? a : 0]; prf[a] = (float)pr[a]; }

  int n = 0, Nn = p.rootN[t], depth = 0, act = 0;
  while (true) {
    const double tN = sT[Nn];
    const float tNf = __double2float_rn(tN);
    // The node's whole child row, then the reciprocals of its visit counts, as two BATCHES of independent,
    // unconditional loads (index clamped past the last action): one memory round trip each.  Loads left inside the
    // per-action conditionals are issued one after the other, each waiting for its own L2 round trip -- measured 13 k
    // cycles per level with 10 actions, against ~1.5 k for the batched form.
    const HotEdge* row = tree + (size_t)n * A;
    HotEdge h[TA];
#pragma unroll
    for (int a = 0; a < TA; ++a) h[a] = row[a < A ? a : 0];
    double rcp[TA];
#pragma unroll
    for (int a = 0; a < TA; ++a) rcp[a] = sR[(h[a].x & 0xffffu) + 1u];
    uint32_t nc[TA], key = 0;
    float s[TA];
#pragma unroll
    for (int a = 0; a < TA; ++a) {
      nc[a] = (uint32_t)kNoChild << 16;
      s[a] = 0.0f;
      if (a < A) {
        nc[a] = h[a].x;
        const float q = __uint_as_float(h[a].y);
        const int cn = (int)(nc[a] & 0xffffu);
        float u;
        if (cn > 0) {
          const double y = div_by_count(tN, cn + 1, rcp[a]);
          u = f32p ? __fmul_rn(prf[a], __double2float_rn(y)) : __double2float_rn(__dmul_rn(pr[a], y));
        } else {
          u = f32p ? __fmul_rn(prf[a], tNf) : __double2float_rn(__dmul_rn(pr[a], tN));
        }
        s[a] = __fadd_rn(q, u);
        key = max(key, f2ord(s[a]));
      }
    }
    const float best = ord2f(key);
    int k = 0;
#pragma unroll
    for (int a = 0; a < TA; ++a) k += (a < A && s[a] == best) ? 1 : 0;
    int r = (k > 1) ? (int)rng.bounded((uint32_t)k) : 0;
    act = 0;
    bool found = false;
#pragma unroll
    for (int a = 0; a < TA; ++a)
      if (!found && a < A && s[a] == best) {
        if (r == 0) { act = a; found = true; }
        --r;
      }
    uint32_t nc_sel = nc[0];
#pragma unroll
    for (int a = 1; a < TA; ++a) nc_sel = (act == a) ? nc[a] : nc_sel;
    pth[depth] = (uint32_t)(n * A + act);
    ++depth;
    if ((nc_sel >> 16) == kNoChild) break;
    n = (int)(nc_sel >> 16);
    Nn = (int)(nc_sel & 0xffffu);
  }
  rng.store(p.rng_pos + t);
  p.leaf_parent[t] = n;
  p.leaf_action[t] = act;
  p.leaf_depth[t] = depth;
  p.src_slot[t] = t * p.max_nodes + n;
  const int c = p.count[t];
  p.dst_slot[t] = t * p.max_nodes + (c < p.max_nodes ? c : p.max_nodes - 1);
  st.depth += (unsigned)depth;
  st.draws += (unsigned)rng.draws;
  st.twists += (unsigned)rng.twists;
  return make_int2(n, act);
}

__device__ __forceinline__ void expand_backup_tree_thread(const PoolDev& p, const int t, const float rew,
                                                          const float val) {
  const int A = p.A;
  const int depth = p.leaf_depth[t];
  const int c = p.count[t];
  if (depth <= 0) return;
  if (c >= p.max_nodes) { atomicOr(p.error, MZ_DEVERR_POOL_FULL); return; }
  const size_t tbase = (size_t)t * p.max_nodes * A;
  HotEdge* tree = p.hot + tbase;
  double* ew = p.ew + tbase;
  float* er = p.er + tbase;
  const uint32_t* pth = p.path + (size_t)t * p.max_nodes;
  for (int a = 0; a < A; ++a) tree[(size_t)c * A + a] = hot_empty();

  double value = (double)val;
  const bool same_pl = p.same_player[t] != 0;
  const bool board = p.board != 0;
  const double discount = p.discount;
  double lo = p.minmax[2 * t], hi = p.minmax[2 * t + 1];
  const double lo0 = lo, hi0 = hi;
  // leaf -> root in chunks of 8 levels: the 8 path entries, then the 8 edge records, are loaded back to back (one
  // memory round trip each instead of one per level), then the value recurrence runs over them in order
  constexpr int CH = 8;
  for (int base = 0; base <= depth; base += CH) {          // i = 0: new leaf ... i = depth: root
    uint32_t eid[CH], rnc[CH];
    double rW[CH];
    float rR[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int level = depth - (base + j);
      eid[j] = level > 0 ? pth[level - 1] : 0u;
    }
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int i = base + j, level = depth - i;
      const bool ld = (i > 0 && level > 0);
      rnc[j] = ld ? tree[eid[j]].x : 0u;
      rW[j] = ld ? ew[eid[j]] : 0.0;
      rR[j] = ld ? er[eid[j]] : 0.0f;
    }
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int i = base + j, level = depth - i;
      if (i > depth) break;
      double W = 0.0, R = 0.0;
      uint32_t N = 0, child = kNoChild;
      if (level > 0) {
        if (i == 0) { R = (double)rew; child = (uint32_t)c; }
        else { W = rW[j]; R = (double)rR[j]; N = rnc[j] & 0xffffu; child = rnc[j] >> 16; }
      } else {
        W = p.rootW[t]; N = (uint32_t)p.rootN[t]; R = p.root_reward[t];
      }
      const bool same = same_pl || ((i & 1) == 0);
      const double Rs = (board && same) ? -R : R;
      const double myval = value;
      value = __dadd_rn(Rs, __dmul_rn(discount, value));
      const double Wn = __dadd_rn(W, same ? myval : -myval);
      const uint32_t Nn = N + 1;
      const double q = __ddiv_rn(Wn, (double)Nn);
      const double mm = __dadd_rn(R, __dmul_rn(discount, board ? -q : q));
      hi = fmax(hi, mm);
      lo = fmin(lo, mm);
      if (level > 0) {
        ew[eid[j]] = Wn;
        if (i == 0) er[eid[j]] = rew;
        tree[eid[j]].x = hot_word(Nn, child);
      } else { p.rootW[t] = Wn; p.rootN[t] = (int)Nn; }
    }
  }
  int* npar = p.node_parent + (size_t)t * p.max_nodes;
  int* nmov = p.node_move + (size_t)t * p.max_nodes;
  p.minmax[2 * t] = lo;
  p.minmax[2 * t + 1] = hi;
  p.count[t] = c + 1;
  npar[c] = p.leaf_parent[t];
  nmov[c] = p.leaf_action[t];
  p.node_value[(size_t)t * p.max_nodes + c] = val;
  p.leaf_depth[t] = 0;
  const bool norm = hi > lo;
  const double range = __dsub_rn(hi, lo);
  const bool moved = __double_as_longlong(lo) != __double_as_longlong(lo0) ||
                     __double_as_longlong(hi) != __double_as_longlong(hi0);
  // child_Q cache: all visited edges (nodes 1..c) when a bound moved, else the path; same chunking
  const int cnt = moved ? c : depth;
  for (int base = 0; base < cnt; base += CH) {
    uint32_t eid[CH], rnc[CH];
    double rW[CH];
    float rR[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int k = base + j;
      eid[j] = k < cnt ? (moved ? (uint32_t)npar[k + 1] * (uint32_t)A + (uint32_t)nmov[k + 1] : pth[k]) : 0u;
    }
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const bool ld = base + j < cnt;
      rnc[j] = ld ? tree[eid[j]].x : 0u;
      rW[j] = ld ? ew[eid[j]] : 0.0;
      rR[j] = ld ? er[eid[j]] : 0.0f;
    }
#pragma unroll
    for (int j = 0; j < CH; ++j)
      if (base + j < cnt)
        tree[eid[j]].y = __float_as_uint(child_q(rW[j], rR[j], rnc[j] & 0xffffu, p.dp, norm, lo, range));
  }
}

}  // namespace mz
