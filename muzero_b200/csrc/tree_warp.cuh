// Warp-per-tree forms of the tree operations: the pUCT descent (select_tree) and expand + backup (expand_backup_tree).
// Shared by the tree kernels of mcts.cu and by the one-launch-per-search kernel of the MLP networks (mlp.cu), where
// warp w of a CTA owns tree w of the CTA's tile for the whole search.
#pragma once
#include "pool.cuh"
#include "tree_thread.cuh"

namespace mz {

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
__device__ __forceinline__ double warp_min_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(kFull, v, o));
  return v;
}

// hot record of child `a` of a row ({child << 16 | N, child_Q}); "no child" past the last action
__device__ __forceinline__ HotEdge load_hot(const HotEdge* row, int a, int A) {
  HotEdge r = hot_empty();
  if (a < A) r = row[a];
  return r;
}

// NCH = number of 32-action chunks held in registers (A <= 32*NCH); NCH == 0: any A, scores staged
// in shared memory.  The register path keeps a node's whole child row, the tree's prior and the
// scores in registers, takes the chosen child's record by shuffle instead of re-reading it, looks
// the pb_c factor and the reciprocals of the visit counts up in shared memory, and loads the row of
// the most-visited child (the one pUCT most often descends into) SPECULATIVELY into registers before
// the scores of the current level are computed: when the descent does go there -- the common case --
// the next level starts without a memory round trip.
// Returns (leaf parent node, action), warp-uniform; also left in the pool's LEAF_* / SRC_SLOT / DST_SLOT / PATH views.
template <int NCH>
__device__ __forceinline__ int2 select_tree(const PoolDev& p, const int t, const int lane, const double* sT,
                                            const double* sR, float* sc, unsigned long long* s_stats) {
  const int A = p.A;
  const HotEdge* tree = p.hot + (size_t)t * p.max_nodes * A;
  const double* __restrict__ P = p.prior + (size_t)t * A;
  const bool f32p = p.f32_prior[t] != 0;
  uint32_t* pth = p.path + (size_t)t * p.max_nodes;

  WarpRng rng;
  rng.load(p.rng_key + (size_t)t * 624, p.rng_pos + t, lane);

  constexpr int NC = NCH > 0 ? NCH : 1;
  double pr[NC];
  float prf[NC];
  uint32_t cur[NC];        // (child << 16 | N) of the current node's edges
  float curq[NC];          // their cached child_Q
  if (NCH > 0) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int a = c * 32 + lane;
      pr[c] = (a < A) ? P[a] : 0.0;
      prf[c] = (float)pr[c];
      const HotEdge h = load_hot(tree, a, A);
      cur[c] = h.x;
      curq[c] = __uint_as_float(h.y);
    }
  }

  int n = 0, Nn = p.rootN[t], depth = 0, act = 0;
  while (true) {
    const double tN = sT[Nn];
    const HotEdge* row = tree + (size_t)n * A;
    uint32_t nc_sel;      // (child << 16 | N) of the chosen edge
    if constexpr (NCH > 0) {
      // most-visited expanded child: its row is the likeliest next one
      uint32_t top = 0;
#pragma unroll
      for (int c = 0; c < NCH; ++c)
        if ((cur[c] >> 16) != kNoChild) top = max(top, (cur[c] << 16) | (cur[c] >> 16));
      top = __reduce_max_sync(kFull, top);
      uint32_t nxt[NC];
      float nxtq[NC];
      const int spec = (top != 0) ? (int)(top & 0xffffu) : -1;
      if (spec >= 0) {
        const HotEdge* srow = tree + (size_t)spec * A;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const HotEdge h = load_hot(srow, c * 32 + lane, A);
          nxt[c] = h.x;
          nxtq[c] = __uint_as_float(h.y);
        }
      }
      // pUCT score = cached child_Q + child_U; the only float64 work left on the chain is y = tN / (cn + 1) for
      // chunks with a visited child and the float64-prior product
      const float tNf = __double2float_rn(tN);
      float s[NC];
      uint32_t key = 0;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int cn = (int)(cur[c] & 0xffffu);
        float u;
        if (__any_sync(kFull, cn > 0)) {
          const double y = div_by_count(tN, cn + 1, __ldg(sR + cn + 1));
          u = f32p ? __fmul_rn(prf[c], __double2float_rn(y)) : __double2float_rn(__dmul_rn(pr[c], y));
        } else {
          // no visited child among these 32 actions (the common case deep in a tree): y = tN / 1 = tN exactly
          u = f32p ? __fmul_rn(prf[c], tNf) : __double2float_rn(__dmul_rn(pr[c], tN));
        }
        s[c] = __fadd_rn(curq[c], u);
        if (c * 32 + lane < A) key = max(key, f2ord(s[c]));
      }
      const float best = ord2f(__reduce_max_sync(kFull, key));
      // ties, ascending action order (np.where(ucb == ucb.max())[0])
      unsigned tb[NC];
      int k = 0;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        tb[c] = __ballot_sync(kFull, (c * 32 + lane < A) && s[c] == best);
        k += __popc(tb[c]);
      }
      int r = (k > 1) ? (int)rng.bounded((uint32_t)k) : 0;
      act = 0;
      bool found = false;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int cnt = __popc(tb[c]);
        if (!found) {
          if (r < cnt) {
            unsigned bb = tb[c];
            for (int q = 0; q < r; ++q) bb &= bb - 1;       // drop the r lowest ties
            act = c * 32 + __ffs(bb) - 1;
            found = true;
          } else {
            r -= cnt;
          }
        }
      }
      // the chosen record from its owner lane: one shuffle per chunk (independent, so one shuffle latency), then a
      // warp-uniform pick -- indexing cur[] with act >> 5 would move the whole row to local memory
      nc_sel = 0;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const uint32_t w = __shfl_sync(kFull, cur[c], act & 31);
        nc_sel = ((act >> 5) == c) ? w : nc_sel;
      }
      const int child = (int)(nc_sel >> 16);
      if (child != (int)kNoChild) {
        if (child == spec) {
#pragma unroll
          for (int c = 0; c < NCH; ++c) { cur[c] = nxt[c]; curq[c] = nxtq[c]; }
        } else {
          const HotEdge* crow = tree + (size_t)child * A;
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            const HotEdge h = load_hot(crow, c * 32 + lane, A);
            cur[c] = h.x;
            curq[c] = __uint_as_float(h.y);
          }
        }
      }
    } else {
      // any A: scores staged in shared memory, same arithmetic (cached child_Q + child_U)
      const float tNf = __double2float_rn(tN);
      float bestl = -INFINITY;
      for (int a = lane; a < A; a += 32) {
        const HotEdge h = row[a];
        const int cn = (int)(h.x & 0xffffu);
        const double pa = P[a];
        float u;
        if (cn > 0) {
          const double y = div_by_count(tN, cn + 1, __ldg(sR + cn + 1));
          u = f32p ? __fmul_rn((float)pa, __double2float_rn(y)) : __double2float_rn(__dmul_rn(pa, y));
        } else {
          u = f32p ? __fmul_rn((float)pa, tNf) : __double2float_rn(__dmul_rn(pa, tN));
        }
        const float s = __fadd_rn(__uint_as_float(h.y), u);
        sc[a] = s;
        bestl = fmaxf(bestl, s);
      }
      const float best = warp_max(bestl);
      __syncwarp();
      int k = 0, first = -1;
      for (int a0 = 0; a0 < A; a0 += 32) {
        const int a = a0 + lane;
        const unsigned b = __ballot_sync(kFull, a < A && sc[a] == best);
        if (first < 0 && b) first = a0 + __ffs(b) - 1;
        k += __popc(b);
      }
      act = first;
      if (k > 1) {
        int r = (int)rng.bounded((uint32_t)k);
        for (int a0 = 0; a0 < A; a0 += 32) {
          const int a = a0 + lane;
          const unsigned b = __ballot_sync(kFull, a < A && sc[a] == best);
          const int c = __popc(b);
          if (r < c) {
            unsigned bb = b;
            for (int q = 0; q < r; ++q) bb &= bb - 1;
            act = a0 + __ffs(bb) - 1;
            break;
          }
          r -= c;
        }
      }
      __syncwarp();
      nc_sel = row[act].x;
    }
    if (lane == 0) pth[depth] = (uint32_t)(n * A + act);
    ++depth;
    if ((nc_sel >> 16) == kNoChild) break;
    n = (int)(nc_sel >> 16);
    Nn = (int)(nc_sel & 0xffffu);
  }
  rng.store(p.rng_pos + t);
  if (lane == 0) {
    p.leaf_parent[t] = n;
    p.leaf_action[t] = act;
    p.leaf_depth[t] = depth;
    p.src_slot[t] = t * p.max_nodes + n;
    const int c = p.count[t];
    p.dst_slot[t] = t * p.max_nodes + (c < p.max_nodes ? c : p.max_nodes - 1);
    // statistics go to the CTA's shared-memory counters; one global atomic per CTA and counter at the end (4096
    // same-address global atomics per launch kept the kernel alive for microseconds after the last descent)
    atomicAdd(s_stats + 0, (unsigned long long)depth);
    atomicAdd(s_stats + 1, 1ULL);
    if (rng.draws) atomicAdd(s_stats + 2, rng.draws);
    if (rng.twists) atomicAdd(s_stats + 3, rng.twists);
  }
  return make_int2(n, act);
}

__device__ __forceinline__ void expand_backup_tree(const PoolDev& p, const int t, const int lane, const float rew,
                                                   const float val) {
  const int A = p.A;
  const int depth = p.leaf_depth[t];
  const int c = p.count[t];
  if (depth <= 0) return;                // no select since the last reset/expand
  if (c >= p.max_nodes) {
    if (lane == 0) atomicOr(p.error, MZ_DEVERR_POOL_FULL);
    return;
  }
  const size_t tbase = (size_t)t * p.max_nodes * A;
  HotEdge* tree = p.hot + tbase;
  double* ew = p.ew + tbase;
  float* er = p.er + tbase;
  const uint32_t* pth = p.path + (size_t)t * p.max_nodes;

  // expand: fresh hot row for the new node (its cold words stay undefined until an edge is first visited)
  HotEdge* crow = tree + (size_t)c * A;
  for (int a = lane; a < A; a += 32) crow[a] = hot_empty();

  double value = (double)val;
  const bool same_pl = p.same_player[t] != 0;
  const bool board = p.board != 0;
  const double discount = p.discount;
  double lo = p.minmax[2 * t], hi = p.minmax[2 * t + 1];
  const double lo0 = lo, hi0 = hi;

  const int total = depth + 1;           // leaf ... root
  // statistics of this lane's edge after the update (last chunk; reused for the child_Q refresh when total <= 32)
  double Wk = 0.0, Rk = 0.0;
  uint32_t Nk = 0, eidk = 0xffffffffu;
  for (int base = 0; base < total; base += 32) {
    const int i = base + lane;           // 0 = new leaf, depth = root
    const bool active = i < total;
    const int level = depth - i;
    double W = 0.0, R = 0.0;
    uint32_t N = 0, child = kNoChild, eid = 0xffffffffu;
    if (active) {
      if (level > 0) {
        eid = pth[level - 1];
        if (i == 0) { R = (double)rew; child = (uint32_t)c; }
        else { const uint32_t nc = tree[eid].x; W = ew[eid]; R = (double)er[eid]; N = nc & 0xffffu; child = nc >> 16; }
      } else {
        W = p.rootW[t]; N = (uint32_t)p.rootN[t]; R = p.root_reward[t];
      }
    }
    // Node.player_id == leaf player  <=>  same parity of depth (players swap every level, mcts.py:379)
    const bool same = same_pl || ((i & 1) == 0);
    const double Rs = (board && same) ? -R : R;
    // serial part of Node.backup: value <- (+-reward) + discount * value, leaf to root
    double myval = 0.0;
    const int cnt = min(32, total - base);
    for (int j = 0; j < cnt; ++j) {
      const double Rj = __shfl_sync(kFull, Rs, j);
      if (lane == j) myval = value;
      value = __dadd_rn(Rj, __dmul_rn(discount, value));
    }
    double mm_hi = -INFINITY, mm_lo = INFINITY;
    if (active) {
      const double Wn = __dadd_rn(W, same ? myval : -myval);
      const uint32_t Nn = N + 1;
      const double q = __ddiv_rn(Wn, (double)Nn);
      const double mm = __dadd_rn(R, __dmul_rn(discount, board ? -q : q));
      mm_hi = mm; mm_lo = mm;
      if (level > 0) {
        ew[eid] = Wn;
        if (i == 0) er[eid] = rew;                          // Node.reward is written once, at expansion
        tree[eid].x = hot_word(Nn, child);
      } else { p.rootW[t] = Wn; p.rootN[t] = (int)Nn; }
      Wk = Wn; Rk = R; Nk = Nn; eidk = (level > 0) ? eid : 0xffffffffu;
    } else {
      eidk = 0xffffffffu;
    }
    hi = fmax(hi, warp_max_d(mm_hi));
    lo = fmin(lo, warp_min_d(mm_lo));
  }
  int* npar = p.node_parent + (size_t)t * p.max_nodes;
  int* nmov = p.node_move + (size_t)t * p.max_nodes;
  if (lane == 0) {
    p.minmax[2 * t] = lo;
    p.minmax[2 * t + 1] = hi;
    p.count[t] = c + 1;
    npar[c] = p.leaf_parent[t];
    nmov[c] = p.leaf_action[t];
    p.node_value[(size_t)t * p.max_nodes + c] = val;
    p.leaf_depth[t] = 0;
  }
  __syncwarp();                          // the records and the node list written above -> every lane
  // child_Q cache for the next descents, under the bounds they will see (the ones just stored)
  {
    const bool norm = hi > lo;
    const double range = __dsub_rn(hi, lo);
    const double dpq = p.dp;
    const bool moved = __double_as_longlong(lo) != __double_as_longlong(lo0) ||
                       __double_as_longlong(hi) != __double_as_longlong(hi0);
    if (moved) {
      // a bound moved: every cached value of this tree is stale.  Visited edges == expanded nodes 1..c
      for (int k = 1 + lane; k <= c; k += 32) {
        const uint32_t e = (uint32_t)npar[k] * (uint32_t)A + (uint32_t)nmov[k];
        tree[e].y = __float_as_uint(child_q(ew[e], er[e], tree[e].x & 0xffffu, dpq, norm, lo, range));
      }
    } else if (total <= 32) {
      // the path's statistics are still in registers (one chunk): no reload
      if (eidk != 0xffffffffu) tree[eidk].y = __float_as_uint(child_q(Wk, (float)Rk, Nk, dpq, norm, lo, range));
    } else {
      for (int k = lane; k < depth; k += 32) {
        const uint32_t e = pth[k];
        tree[e].y = __float_as_uint(child_q(ew[e], er[e], tree[e].x & 0xffffu, dpq, norm, lo, range));
      }
    }
  }
}

__device__ __forceinline__ void flush_stats(const PoolDev& p, const unsigned long long* s_stats) {
  __syncthreads();
  if (threadIdx.x < 4 && s_stats[threadIdx.x]) atomicAdd(p.stats + threadIdx.x, s_stats[threadIdx.x]);
}

}  // namespace mz
