"""Device-resident callers of the search path.

Reference interfaces mirrored here (michaelnny/muzero):
  muzero/games/env.py:38-154,242-302      BoardGameEnv.reset / step / observation / actions_mask
  muzero/games/gomoku.py:72-116           GomokuEnv.is_current_player_won   (tictactoe.py:33-77 likewise)
  muzero/pipeline.py:632-671              compute_n_step_target
  muzero/pipeline.py:674-706              compute_mc_return_target
  muzero/pipeline.py:709-767              make_unroll_sequence
  muzero/pipeline.py:91-165               the self-play loop of run_self_play (board games)

Everything computes in the CUDA kernels of csrc/selfplay.cu behind the C ABI; this module only owns tensors and
sequences launches.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List, NamedTuple, Optional

import numpy as np
import torch

from . import _lib


class BatchedBoardEnv:
    """G independent games of one board game on the GPU (``GomokuEnv`` / ``TicTacToeEnv`` semantics).

    ``obs`` float32 [G, 2*stack+1, N, N] is the observation of the side to move, ``actions_mask`` uint8 [G, A] the
    legal actions (A = N*N + 1, the last action resigns), ``current_player`` int32 [G] (1 black, 2 white).
    """

    def __init__(self, num_games: int, board_size: int, num_to_win: int, stack_history: int,
                 device='cuda') -> None:
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('BatchedBoardEnv lives on a CUDA device (no CPU fallback)')
        if self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        self.G, self.N, self.k, self.stack = int(num_games), int(board_size), int(num_to_win), int(stack_history)
        self.num_actions = self.N * self.N + 1
        self.resign_action = self.num_actions - 1
        lib = _lib.lib()
        n = C.c_size_t()
        _lib.check(lib.mz_env_arena_bytes(self.G, self.N, self.stack, C.byref(n)))
        with torch.cuda.device(self.device):
            self.arena = torch.zeros(n.value + 256, dtype=torch.uint8, device=self.device)
            base = (self.arena.data_ptr() + 255) // 256 * 256
            self.handle = C.c_void_p()
            _lib.check(lib.mz_env_create(self.G, self.N, self.stack, self.k, base, n.value, C.byref(self.handle)))
        G, nn = self.G, self.N * self.N
        self.board = self._view(0, torch.int8).view(G, self.N, self.N)
        self.history = self._view(1, torch.int8).view(G, 2, self.stack, self.N, self.N)
        self.actions_mask = self._view(2, torch.uint8).view(G, self.num_actions)
        self.current_player = self._view(3, torch.int32)
        self.steps = self._view(4, torch.int32)
        self.winner = self._view(5, torch.int32)
        self.done = self._view(6, torch.uint8)
        self._error = self._view(7, torch.int32)
        self.obs = torch.zeros((G, 2 * self.stack + 1, self.N, self.N), dtype=torch.float32, device=self.device)
        self.reward = torch.zeros(G, dtype=torch.float64, device=self.device)
        self.step_done = torch.zeros(G, dtype=torch.uint8, device=self.device)
        self.mover = torch.zeros(G, dtype=torch.int32, device=self.device)
        self.reset()

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                _lib.lib().mz_env_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def _view(self, which: int, dtype) -> torch.Tensor:
        p, n = C.c_void_p(), C.c_size_t()
        _lib.check(_lib.lib().mz_env_view(self.handle, which, C.byref(p), C.byref(n)))
        off = p.value - self.arena.data_ptr()
        return self.arena[off:off + n.value].view(dtype)

    def reset(self, which: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Reset every game, or those with ``which[g] != 0``.  Returns ``obs``."""
        if which is not None:
            which = which.to(device=self.device, dtype=torch.uint8).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().mz_env_reset(self.handle, _lib.ptr(which), _lib.ptr(self.obs), _lib.current_stream()))
        return self.obs

    def step(self, action: torch.Tensor):
        """One move per unfinished game.  Returns (obs, reward f64[G], done u8[G], mover i32[G]) — tensors this
        object owns and overwrites at the next step."""
        action = action.to(device=self.device, dtype=torch.int32).contiguous()
        assert action.shape == (self.G,)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().mz_env_step(self.handle, _lib.ptr(action), _lib.ptr(self.reward),
                                              _lib.ptr(self.step_done), _lib.ptr(self.mover), _lib.ptr(self.obs),
                                              _lib.current_stream()))
        return self.obs, self.reward, self.step_done, self.mover

    def check_errors(self) -> None:
        """Synchronising read of the sticky error word (the reference raises ValueError at the offending step)."""
        e = int(self._error.cpu()[0])
        if e:
            self._error.zero_()
            raise ValueError('Invalid action submitted to BatchedBoardEnv.step (illegal or already taken)')


def _pow_table(discount: float, n: int, device) -> torch.Tensor:
    return torch.tensor([discount ** i for i in range(n + 1)], dtype=torch.float64, device=device)


def n_step_targets(lengths: torch.Tensor, rewards: torch.Tensor, root_values: torch.Tensor, td_steps: int,
                   discount: float):
    """``compute_n_step_target`` for G padded trajectories [G, Tmax] (float64).  Returns (targets, priorities)."""
    G, T = rewards.shape
    targets, prio = torch.empty_like(rewards), torch.empty_like(rewards)
    tab = _pow_table(float(discount), int(td_steps), rewards.device)
    with torch.cuda.device(rewards.device):
        _lib.check(_lib.lib().mz_targets_nstep(G, T, _lib.ptr(lengths), _lib.ptr(rewards), _lib.ptr(root_values),
                                               int(td_steps), _lib.ptr(tab), _lib.ptr(targets), _lib.ptr(prio),
                                               _lib.current_stream()))
    return targets, prio


def mc_return_targets(lengths: torch.Tensor, rewards: torch.Tensor, player_ids: torch.Tensor,
                      root_values: torch.Tensor):
    """``compute_mc_return_target`` (board games) + priorities for G padded trajectories."""
    G, T = rewards.shape
    targets, prio = torch.empty_like(rewards), torch.empty_like(rewards)
    with torch.cuda.device(rewards.device):
        _lib.check(_lib.lib().mz_targets_mc(G, T, _lib.ptr(lengths), _lib.ptr(rewards), _lib.ptr(player_ids),
                                            _lib.ptr(root_values), _lib.ptr(targets), _lib.ptr(prio),
                                            _lib.current_stream()))
    return targets, prio


def unroll_sequences(lengths: torch.Tensor, actions: torch.Tensor, rewards: torch.Tensor, targets: torch.Tensor,
                     pi: torch.Tensor, unroll_steps: int):
    """``make_unroll_sequence`` for G padded trajectories.  Returns (action i32 [G,T,K], reward f32 [G,T,K],
    value f32 [G,T,K], pi f32 [G,T,K,A], valid bool [G,T])."""
    G, T = rewards.shape
    A, K, dev = pi.shape[-1], int(unroll_steps), rewards.device
    oa = torch.zeros((G, T, K), dtype=torch.int32, device=dev)
    orw = torch.zeros((G, T, K), dtype=torch.float32, device=dev)
    ov = torch.zeros((G, T, K), dtype=torch.float32, device=dev)
    op = torch.zeros((G, T, K, A), dtype=torch.float32, device=dev)
    valid = torch.zeros((G, T), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().mz_unroll_sequences(G, T, K, A, _lib.ptr(lengths), _lib.ptr(actions), _lib.ptr(rewards),
                                                  _lib.ptr(targets), _lib.ptr(pi), _lib.ptr(oa), _lib.ptr(orw),
                                                  _lib.ptr(ov), _lib.ptr(op), _lib.ptr(valid), _lib.current_stream()))
    return oa, orw, ov, op, valid.bool()


class Samples(NamedTuple):
    """Training samples of finished games, the fields of the reference's ``Transition`` (replay.py) + priority."""
    state: torch.Tensor       # int8  [n, C, N, N]
    action: torch.Tensor      # int32 [n, K]
    reward: torch.Tensor      # f32   [n, K]
    value: torch.Tensor       # f32   [n, K]
    pi_prob: torch.Tensor     # f32   [n, K, A]
    priority: torch.Tensor    # f64   [n]


class BoardSelfPlay:
    """``run_self_play`` (pipeline.py:91-165) for G concurrent board games, everything on the GPU: batched search ->
    batched env step -> trajectory buffers -> MC-return targets and unroll windows when a game ends -> the slot
    restarts at once.  One small device->host read per move (which games ended)."""

    def __init__(self, network, config, env: BatchedBoardEnv, train_steps: int = 0, seed: int = 0) -> None:
        assert config.is_board_game
        from .mcts import _plan_for
        self.net, self.cfg, self.env = network, config, env
        self._plan = _plan_for(network, config, env.G)
        self._plan.pool.seed(seed + np.arange(env.G))     # one MT19937 stream per game slot
        G, A, dev = env.G, env.num_actions, env.device
        self.Tmax = env.N * env.N + 1
        self.train_steps = train_steps
        C_, N = 2 * env.stack + 1, env.N
        self.t_obs = torch.zeros((G, self.Tmax, C_, N, N), dtype=torch.int8, device=dev)
        self.t_action = torch.zeros((G, self.Tmax), dtype=torch.int32, device=dev)
        self.t_reward = torch.zeros((G, self.Tmax), dtype=torch.float64, device=dev)
        self.t_root = torch.zeros((G, self.Tmax), dtype=torch.float64, device=dev)
        self.t_player = torch.zeros((G, self.Tmax), dtype=torch.int32, device=dev)
        self.t_pi = torch.zeros((G, self.Tmax, A), dtype=torch.float32, device=dev)
        self.steps = np.zeros(G, dtype=np.int64)          # host mirror of env.steps (drives temperature and players)
        self._rows = torch.arange(G, device=dev)
        self.games_finished = 0
        self.moves_played = 0

    def play_move(self) -> Optional[Samples]:
        """Search + one move in every game; returns the samples of the games that just ended (or None)."""
        import muzero_b200 as mz
        env, cfg, G = self.env, self.cfg, self.env.G
        # config.visit_softmax_temperature_fn(steps, train_steps) per game (pipeline.py:100): one call per distinct
        # move number instead of one per game
        uniq, inv = np.unique(self.steps, return_inverse=True)
        temps = np.array([cfg.visit_softmax_temperature_fn(int(s), self.train_steps) for s in uniq], np.float64)[inv]
        cur = (1 + (self.steps % 2)).astype(np.int32)                   # black moves first, players alternate
        # plan=: the plan this object seeded (the module-level plan cache may have evicted and rebuilt its entry)
        action, pi, root_value = mz.uct_search_batch(env.obs, self.net, cfg, temps, env.actions_mask, cur, 3 - cur,
                                                     plan=self._plan)
        # the reference's actor dies with ValueError('probabilities contain NaN') when every visit of a search went
        # below an illegal first pick (mcts.py:404); same here, before the bad action reaches the environment
        self._plan.pool.check_errors()
        t = torch.from_numpy(self.steps).to(env.device)
        self.t_obs[self._rows, t] = env.obs.to(torch.int8)
        self.t_action[self._rows, t] = action
        self.t_root[self._rows, t] = root_value
        self.t_pi[self._rows, t] = pi.to(torch.float32)
        _, reward, done, mover = env.step(action)
        self.t_reward[self._rows, t] = reward
        self.t_player[self._rows, t] = mover
        self.steps += 1
        self.moves_played += G
        done_h = done.cpu().numpy().astype(bool)
        if not done_h.any():
            return None
        lengths = torch.from_numpy(np.where(done_h, self.steps, 0).astype(np.int32)).to(env.device)
        targets, prio = mc_return_targets(lengths, self.t_reward, self.t_player, self.t_root)
        sa, sr, sv, sp, valid = unroll_sequences(lengths, self.t_action, self.t_reward, targets, self.t_pi,
                                                 cfg.unroll_steps)
        out = Samples(self.t_obs[valid], sa[valid], sr[valid], sv[valid], sp[valid], prio[valid])
        env.reset(done)
        self.steps[done_h] = 0
        self.games_finished += int(done_h.sum())
        return out
