// Callers of the search hot path, device-resident:
//   * batched board-game environments: BoardGameEnv.step / reset / observation (games/env.py:117-154,
//     242-271, 295-302) with the last-move win check of games/gomoku.py:72-116 / games/tictactoe.py:33-77;
//   * trajectory -> training targets: compute_n_step_target / compute_mc_return_target /
//     make_unroll_sequence (pipeline.py:632-767) and the priorities of pipeline.py:128,152.
// Plain integer / float64 work, one thread per game or per (game, step); no tensor cores here.
// float64 arithmetic is issued as separate IEEE operations in the reference's order (no FMA contraction),
// so targets and priorities are bit-identical to the reference's Python floats.
#include "common.cuh"

namespace mz {

// ---------------------------------------------------------------------------
// board environments
// ---------------------------------------------------------------------------
struct EnvDev {
  int G, N, stack, num_to_win, A;      // games, board size, history planes per player, stones in a row, actions = N*N + 1
  int8_t* board;        // [G][N*N] 0 empty, 1 black, 2 white (games/env.py:69-70)
  int8_t* hist;         // [G][2][stack][N*N] per-player FIFO of own-stone planes, most recent first (env.py:295-302)
  uint8_t* mask;        // [G][A] legal actions (env.py:83)
  int32_t* player;      // [G] player to move: 1 black, 2 white (env.py:89)
  int32_t* steps;       // [G]
  int32_t* winner;      // [G] 0 none, 1, 2
  uint8_t* done;        // [G]
  int32_t* error;       // sticky: bit 0 illegal action, bit 1 step after game over
};

__device__ __forceinline__ int count_dir(const int8_t* b, int N, int r, int c, int dr, int dc, int color) {
  // count_same_color_stones (games/gomoku.py): the start stone plus same-coloured stones along (dr, dc)
  int n = 0;
  while (r >= 0 && r < N && c >= 0 && c < N && b[r * N + c] == color) { ++n; r += dr; c += dc; }
  return n;
}

// observation(): [X_t, Y_t, X_t-1, Y_t-1, ..., C] from the side to move (env.py:242-271), written as float32
__device__ void write_observation(const EnvDev& e, int g, float* obs) {
  const int nn = e.N * e.N;
  const int cur = e.player[g], opp = 3 - cur;
  const int8_t* hc = e.hist + ((size_t)g * 2 + (cur - 1)) * e.stack * nn;
  const int8_t* ho = e.hist + ((size_t)g * 2 + (opp - 1)) * e.stack * nn;
  float* o = obs + (size_t)g * (2 * e.stack + 1) * nn;
  for (int t = 0; t < e.stack; ++t)
    for (int i = 0; i < nn; ++i) {
      o[(2 * t) * nn + i] = (float)hc[t * nn + i];
      o[(2 * t + 1) * nn + i] = (float)ho[t * nn + i];
    }
  const float colour = cur == 1 ? 1.0f : 0.0f;
  for (int i = 0; i < nn; ++i) o[2 * e.stack * nn + i] = colour;
}

__global__ void env_reset_kernel(EnvDev e, const uint8_t* __restrict__ which, float* __restrict__ obs) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= e.G || (which != nullptr && which[g] == 0)) return;
  const int nn = e.N * e.N;
  for (int i = 0; i < nn; ++i) e.board[(size_t)g * nn + i] = 0;
  for (int i = 0; i < 2 * e.stack * nn; ++i) e.hist[(size_t)g * 2 * e.stack * nn + i] = 0;
  for (int a = 0; a < e.A; ++a) e.mask[(size_t)g * e.A + a] = 1;
  e.player[g] = 1; e.steps[g] = 0; e.winner[g] = 0; e.done[g] = 0;
  if (obs) write_observation(e, g, obs);
}

// step(action) for every game that is not done (env.py:117-154).  reward / done of THIS move; the observation
// is the next position seen by the side to move next.  Finished games are left untouched (reward 0, done 1).
__global__ void env_step_kernel(EnvDev e, const int32_t* __restrict__ action, double* __restrict__ reward,
                                uint8_t* __restrict__ done_out, int32_t* __restrict__ mover, float* __restrict__ obs) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= e.G) return;
  const int nn = e.N * e.N;
  if (e.done[g]) {
    if (reward) reward[g] = 0.0;
    if (done_out) done_out[g] = 1;
    if (mover) mover[g] = e.player[g];
    return;
  }
  const int a = action[g];
  const int cur = e.player[g];
  if (mover) mover[g] = cur;
  if (a < 0 || a >= e.A || e.mask[(size_t)g * e.A + a] == 0) {      // ValueError in the reference (env.py:119-122)
    atomicOr(e.error, 1);
    if (reward) reward[g] = 0.0;
    if (done_out) done_out[g] = 0;
    return;
  }
  double r = 0.0;
  e.mask[(size_t)g * e.A + a] = 0;
  int8_t* b = e.board + (size_t)g * nn;
  if (a == e.A - 1) {                    // resign: always a loss for the mover (env.py:134-136)
    r = -1.0;
    e.winner[g] = 3 - cur;
  } else {
    const int row = a / e.N, col = a % e.N;
    b[a] = (int8_t)cur;                  // colour == player id (black 1, white 2)
    // _update_feature_planes: push the mover's own stones on ITS queue (env.py:295-302)
    int8_t* h = e.hist + ((size_t)g * 2 + (cur - 1)) * e.stack * nn;
    for (int t = e.stack - 1; t > 0; --t)
      for (int i = 0; i < nn; ++i) h[t * nn + i] = h[(t - 1) * nn + i];
    for (int i = 0; i < nn; ++i) h[i] = b[i] == cur ? 1 : 0;
    // is_current_player_won: only once both sides could have num_to_win stones (gomoku.py:75-77), from the last move
    if (e.steps[g] >= (e.num_to_win - 1) * 2) {
      const int k = e.num_to_win;
      const bool won = count_dir(b, e.N, row, col, 0, -1, cur) + count_dir(b, e.N, row, col, 0, 1, cur) - 1 >= k ||
                       count_dir(b, e.N, row, col, -1, 0, cur) + count_dir(b, e.N, row, col, 1, 0, cur) - 1 >= k ||
                       count_dir(b, e.N, row, col, -1, -1, cur) + count_dir(b, e.N, row, col, 1, 1, cur) - 1 >= k ||
                       count_dir(b, e.N, row, col, -1, 1, cur) + count_dir(b, e.N, row, col, 1, -1, cur) - 1 >= k;
      if (won) { r = 1.0; e.winner[g] = cur; }
    }
  }
  bool over = e.winner[g] != 0;
  if (!over) {                           // is_board_full (env.py:344-346)
    over = true;
    for (int i = 0; i < nn; ++i) if (b[i] == 0) { over = false; break; }
  }
  e.done[g] = over ? 1 : 0;
  if (!over) e.player[g] = 3 - cur;      // the next player only moves on when the game goes on (env.py:149-151)
  e.steps[g] += 1;
  if (reward) reward[g] = r;
  if (done_out) done_out[g] = over ? 1 : 0;
  if (obs) write_observation(e, g, obs);
}

// ---------------------------------------------------------------------------
// trajectory -> targets
// ---------------------------------------------------------------------------
// compute_n_step_target (pipeline.py:632-671): z_t = sum_{i<n} discount**i * r[t+i] + discount**n * v[t+n],
// zeros past the end; discount**i comes from a host table evaluated with CPython's float pow.
__global__ void nstep_target_kernel(int G, int Tmax, const int32_t* __restrict__ len, const double* __restrict__ rewards,
                                    const double* __restrict__ root_values, int n, const double* __restrict__ powtab,
                                    double* __restrict__ targets, double* __restrict__ priorities) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G * Tmax) return;
  const int g = i / Tmax, t = i % Tmax, T = len[g];
  if (t >= T) { targets[i] = 0.0; if (priorities) priorities[i] = 0.0; return; }
  const double* r = rewards + (size_t)g * Tmax;
  const double* v = root_values + (size_t)g * Tmax;
  // sum([...]) of CPython >= 3.12 is Neumaier-compensated (bltinmodule.c builtin_sum_impl): 0 + x0 exactly, then
  // t = f + x; c += |f| >= |x| ? (f - t) + x : (x - t) + f; f = t; and f += c at the end when c is finite and non-zero
  double value = 0.0;
  if (n > 0) {
    double f = __dadd_rn(0.0, __dmul_rn(powtab[0], r[t])), c = 0.0;      // int 0 + first item
    for (int k = 1; k < n; ++k) {
      const double x = __dmul_rn(powtab[k], t + k < T ? r[t + k] : 0.0);
      const double s = __dadd_rn(f, x);
      if (fabs(f) >= fabs(x)) c = __dadd_rn(c, __dadd_rn(__dsub_rn(f, s), x));
      else c = __dadd_rn(c, __dadd_rn(__dsub_rn(x, s), f));
      f = s;
    }
    if (c != 0.0 && isfinite(c)) f = __dadd_rn(f, c);
    value = f;
  }
  value = __dadd_rn(value, __dmul_rn(powtab[n], t + n < T ? v[t + n] : 0.0));
  targets[i] = value;
  if (priorities) priorities[i] = fabs(__dsub_rn(v[t], value));      // pipeline.py:128,152
}

// compute_mc_return_target (pipeline.py:674-706): +-final reward by who moved, board games only
__global__ void mc_target_kernel(int G, int Tmax, const int32_t* __restrict__ len, const double* __restrict__ rewards,
                                 const int32_t* __restrict__ player_ids, const double* __restrict__ root_values,
                                 double* __restrict__ targets, double* __restrict__ priorities) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G * Tmax) return;
  const int g = i / Tmax, t = i % Tmax, T = len[g];
  double value = 0.0;
  if (t < T) {
    const double fr = rewards[(size_t)g * Tmax + T - 1];
    const int fp = player_ids[(size_t)g * Tmax + T - 1];
    if (fr != 0.0) value = player_ids[i] == fp ? fr : -fr;
  }
  targets[i] = value;
  if (priorities) priorities[i] = t < T ? fabs(__dsub_rn(root_values[i], value)) : 0.0;
}

// make_unroll_sequence (pipeline.py:709-767): K-step windows with absorbing padding (action 0, reward 0, value 0,
// uniform policy) past the end of the game.  Actions stay int32 (the reference's int8 cast wraps for A > 127).
__global__ void unroll_kernel(int G, int Tmax, int K, int A, const int32_t* __restrict__ len,
                              const int32_t* __restrict__ actions, const double* __restrict__ rewards,
                              const double* __restrict__ targets, const float* __restrict__ pi, float uniform,
                              int32_t* __restrict__ out_action, float* __restrict__ out_reward,
                              float* __restrict__ out_value, float* __restrict__ out_pi, uint8_t* __restrict__ valid) {
  const int i = blockIdx.x;                // (game, step)
  const int g = i / Tmax, t = i % Tmax, T = len[g];
  if (threadIdx.x == 0 && valid) valid[i] = t < T ? 1 : 0;
  if (t >= T) return;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const bool in = t + k < T;
    const size_t s = (size_t)g * Tmax + t + k;
    out_action[(size_t)i * K + k] = in ? actions[s] : 0;
    out_reward[(size_t)i * K + k] = in ? (float)rewards[s] : 0.0f;
    out_value[(size_t)i * K + k] = in ? (float)targets[s] : 0.0f;
  }
  for (int j = threadIdx.x; j < K * A; j += blockDim.x) {
    const int k = j / A, a = j % A;
    out_pi[(size_t)i * K * A + j] = t + k < T ? pi[((size_t)g * Tmax + t + k) * A + a] : uniform;
  }
}

}  // namespace mz

using namespace mz;

struct mz_env {
  EnvDev d;
};

extern "C" int mz_env_arena_bytes(int32_t games, int32_t board_size, int32_t stack_history, size_t* bytes) {
  MZ_CHECK_ARG(bytes && games > 0 && board_size > 0 && board_size <= 32 && stack_history > 0, "bad environment shape");
  const size_t nn = (size_t)board_size * board_size, G = games;
  *bytes = align_up(G * nn, 256) + align_up(G * 2 * stack_history * nn, 256) + align_up(G * (nn + 1), 256) +
           4 * align_up(G * 4, 256) + align_up(G, 256) + 256;
  return MZ_OK;
}

extern "C" int mz_env_create(int32_t games, int32_t board_size, int32_t stack_history, int32_t num_to_win,
                             void* arena_dev, size_t arena_bytes, mz_env** out) {
  size_t need;
  int rc = mz_env_arena_bytes(games, board_size, stack_history, &need);
  if (rc) return rc;
  MZ_CHECK_ARG(arena_dev && out && num_to_win > 0, "NULL argument");
  MZ_CHECK_ARG(((uintptr_t)arena_dev & 255) == 0, "arena must be 256-byte aligned");
  if (arena_bytes < need) { set_error("env arena too small: %zu < %zu", arena_bytes, need); return MZ_ENOMEM; }
  mz_env* e = new mz_env();
  EnvDev& d = e->d;
  const size_t nn = (size_t)board_size * board_size, G = games;
  d.G = games; d.N = board_size; d.stack = stack_history; d.num_to_win = num_to_win; d.A = (int)nn + 1;
  char* p = (char*)arena_dev;
  auto take = [&](size_t b) { char* r = p; p += align_up(b, 256); return r; };
  d.board = (int8_t*)take(G * nn);
  d.hist = (int8_t*)take(G * 2 * stack_history * nn);
  d.mask = (uint8_t*)take(G * (nn + 1));
  d.player = (int32_t*)take(G * 4);
  d.steps = (int32_t*)take(G * 4);
  d.winner = (int32_t*)take(G * 4);
  d.error = (int32_t*)take(G * 4);
  d.done = (uint8_t*)take(G);
  MZ_CUDA(cudaMemset(d.error, 0, 4));
  *out = e;
  return MZ_OK;
}

extern "C" int mz_env_destroy(mz_env* env) {
  delete env;
  return MZ_OK;
}

extern "C" int mz_env_view(mz_env* env, int32_t which, void** ptr, size_t* bytes) {
  MZ_CHECK_ARG(env && ptr && bytes, "NULL argument");
  const EnvDev& d = env->d;
  const size_t nn = (size_t)d.N * d.N, G = d.G;
  switch (which) {
    case 0: *ptr = d.board; *bytes = G * nn; break;
    case 1: *ptr = d.hist; *bytes = G * 2 * d.stack * nn; break;
    case 2: *ptr = d.mask; *bytes = G * d.A; break;
    case 3: *ptr = d.player; *bytes = G * 4; break;
    case 4: *ptr = d.steps; *bytes = G * 4; break;
    case 5: *ptr = d.winner; *bytes = G * 4; break;
    case 6: *ptr = d.done; *bytes = G; break;
    case 7: *ptr = d.error; *bytes = 4; break;
    default: set_error("unknown env view %d", which); return MZ_EINVAL;
  }
  return MZ_OK;
}

extern "C" int mz_env_reset(mz_env* env, const uint8_t* which, float* obs, mz_stream stream) {
  MZ_CHECK_ARG(env, "NULL argument");
  env_reset_kernel<<<(env->d.G + 127) / 128, 128, 0, (cudaStream_t)stream>>>(env->d, which, obs);
  MZ_LAUNCH_CHECK("env_reset_kernel");
  return MZ_OK;
}

extern "C" int mz_env_step(mz_env* env, const int32_t* action, double* reward, uint8_t* done, int32_t* mover,
                           float* obs, mz_stream stream) {
  MZ_CHECK_ARG(env && action, "NULL argument");
  env_step_kernel<<<(env->d.G + 127) / 128, 128, 0, (cudaStream_t)stream>>>(env->d, action, reward, done, mover, obs);
  MZ_LAUNCH_CHECK("env_step_kernel");
  return MZ_OK;
}

extern "C" int mz_targets_nstep(int32_t games, int32_t max_len, const int32_t* lengths, const double* rewards,
                                const double* root_values, int32_t td_steps, const double* pow_table_dev,
                                double* targets, double* priorities, mz_stream stream) {
  MZ_CHECK_ARG(lengths && rewards && root_values && pow_table_dev && targets, "NULL argument");
  MZ_CHECK_ARG(games > 0 && max_len > 0 && td_steps >= 0, "bad sizes");
  const int n = games * max_len;
  nstep_target_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(games, max_len, lengths, rewards, root_values,
                                                                        td_steps, pow_table_dev, targets, priorities);
  MZ_LAUNCH_CHECK("nstep_target_kernel");
  return MZ_OK;
}

extern "C" int mz_targets_mc(int32_t games, int32_t max_len, const int32_t* lengths, const double* rewards,
                             const int32_t* player_ids, const double* root_values, double* targets,
                             double* priorities, mz_stream stream) {
  MZ_CHECK_ARG(lengths && rewards && player_ids && targets, "NULL argument");
  MZ_CHECK_ARG(priorities == nullptr || root_values != nullptr, "priorities need root_values");
  MZ_CHECK_ARG(games > 0 && max_len > 0, "bad sizes");
  const int n = games * max_len;
  mc_target_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(games, max_len, lengths, rewards, player_ids,
                                                                     root_values, targets, priorities);
  MZ_LAUNCH_CHECK("mc_target_kernel");
  return MZ_OK;
}

extern "C" int mz_unroll_sequences(int32_t games, int32_t max_len, int32_t unroll_steps, int32_t num_actions,
                                   const int32_t* lengths, const int32_t* actions, const double* rewards,
                                   const double* targets, const float* pi, int32_t* out_action, float* out_reward,
                                   float* out_value, float* out_pi, uint8_t* valid, mz_stream stream) {
  MZ_CHECK_ARG(lengths && actions && rewards && targets && pi && out_action && out_reward && out_value && out_pi,
               "NULL argument");
  MZ_CHECK_ARG(games > 0 && max_len > 0 && unroll_steps > 0 && num_actions > 0, "bad sizes");
  const float uniform = (float)(1.0 / (double)num_actions);       // np.ones_like(pi) / len(pi), cast to float32
  unroll_kernel<<<games * max_len, 128, 0, (cudaStream_t)stream>>>(games, max_len, unroll_steps, num_actions, lengths,
                                                                  actions, rewards, targets, pi, uniform, out_action,
                                                                  out_reward, out_value, out_pi, valid);
  MZ_LAUNCH_CHECK("unroll_kernel");
  return MZ_OK;
}
