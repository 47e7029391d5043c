"""CPU restatement of the callers of the search path (TEST INFRASTRUCTURE: only tests/, smoke() and bench.py's
cpu_baseline may import this; the product never does).

* ``BoardEnvOracle``: ``BoardGameEnv`` of michaelnny/muzero (muzero/games/env.py:38-154, 242-302, 344-353) with the
  last-move win check of muzero/games/gomoku.py:72-116 / muzero/games/tictactoe.py:33-77, as plain numpy arrays.
* ``n_step_target`` / ``mc_return_target`` / ``unroll_sequences``: muzero/pipeline.py:632-671, 674-706, 709-767.

Pinned: tests/golden/make_golden_selfplay.py runs the UNMODIFIED reference classes / functions (imported from
/root/reference with gym / snappy stand-ins) on random games and trajectories and stores what they return;
tests/test_selfplay.py replays those recordings through this module bit-for-bit.
"""
from __future__ import annotations

import numpy as np


class BoardEnvOracle:
    def __init__(self, board_size: int, num_to_win: int, stack_history: int) -> None:
        self.N, self.k, self.stack = board_size, num_to_win, stack_history
        self.A = board_size * board_size + 1                         # env.py:76 (resign enabled)
        self.resign = self.A - 1                                     # env.py:86
        self.reset()

    def reset(self) -> np.ndarray:                                   # env.py:100-115
        n = self.N
        self.board = np.zeros((n, n), dtype=np.int8)
        self.mask = np.ones(self.A, dtype=bool)
        self.player, self.steps, self.winner = 1, 0, None
        self.hist = {1: [np.zeros((n, n), dtype=np.int8) for _ in range(self.stack)],
                     2: [np.zeros((n, n), dtype=np.int8) for _ in range(self.stack)]}
        return self.observation()

    @property
    def opponent(self) -> int:
        return 3 - self.player

    @property
    def done(self) -> bool:                                          # env.py:348-353
        return self.winner is not None or bool(np.all(self.board != 0))

    def _run(self, r, c, dr, dc, colour) -> int:                     # count_same_color_stones, gomoku.py
        n = 0
        while 0 <= r < self.N and 0 <= c < self.N and self.board[r, c] == colour:
            n, r, c = n + 1, r + dr, c + dc
        return n

    def _won(self, r, c, colour) -> bool:                            # gomoku.py:72-116
        if self.steps < (self.k - 1) * 2:
            return False
        for (a, b), (c2, d) in (((0, -1), (0, 1)), ((-1, 0), (1, 0)), ((-1, -1), (1, 1)), ((-1, 1), (1, -1))):
            if self._run(r, c, a, b, colour) + self._run(r, c, c2, d, colour) - 1 >= self.k:
                return True
        return False

    def step(self, action: int):                                     # env.py:117-154
        if not 0 <= action <= self.A - 1:
            raise ValueError('Invalid action')
        if not self.mask[action]:
            raise ValueError('Invalid action. The action has alread been taken.')
        if self.done:
            raise RuntimeError('Game is over, call reset before using step method.')
        reward = 0.0
        self.mask[action] = False
        if action == self.resign:
            reward, self.winner = -1.0, self.opponent
        else:
            r, c = divmod(action, self.N)
            self.board[r, c] = self.player                           # colour == player id (env.py:66-70)
            q = self.hist[self.player]                               # env.py:295-302
            q.insert(0, (self.board == self.player).astype(np.int8))
            q.pop()
            if self._won(r, c, self.player):
                reward, self.winner = 1.0, self.player
        done = self.done
        if not done:
            self.player = self.opponent
        self.steps += 1
        return self.observation(), reward, done

    def observation(self) -> np.ndarray:                             # env.py:242-271
        planes = []
        for t in range(self.stack):
            planes.append(self.hist[self.player][t])
            planes.append(self.hist[self.opponent][t])
        colour = np.full((1, self.N, self.N), 1 if self.player == 1 else 0, dtype=np.int8)
        return np.concatenate([np.array(planes, dtype=np.int8), colour], axis=0)


def _py312_sum(items):
    """CPython >= 3.12 ``sum()`` of floats: Neumaier compensated summation (Python/bltinmodule.c, builtin_sum_impl).
    The parity target is the reference executed by THIS container's interpreter (3.12), so the compensation is part
    of what ``compute_n_step_target`` computes."""
    if not items:
        return 0
    f, c = 0 + items[0], 0.0
    for x in items[1:]:
        t = f + x
        if abs(f) >= abs(x):
            c += (f - t) + x
        else:
            c += (x - t) + f
        f = t
    if c and np.isfinite(c):
        f += c
    return f


def n_step_target(rewards, root_values, td_steps: int, discount: float):
    """pipeline.py:632-671, Python floats, same operation order."""
    if len(rewards) != len(root_values):
        raise ValueError('Arguments `rewards` and `root_values` don have the same length.')
    T = len(rewards)
    r = [float(x) for x in rewards] + [0] * td_steps
    v = [float(x) for x in root_values] + [0] * td_steps
    out = []
    for t in range(T):
        value = _py312_sum([discount ** i * r[t + i] for i in range(td_steps)])
        value = value + discount ** td_steps * v[t + td_steps]
        out.append(value)
    return out


def mc_return_target(rewards, player_ids):
    """pipeline.py:674-706."""
    if len(rewards) != len(player_ids):
        raise ValueError('Arguments `rewards` and `player_ids` don have the same length.')
    T = len(rewards)
    out = [0.0] * T
    if rewards[-1] != 0.0:
        for t in range(T):
            out[t] = rewards[-1] if player_ids[t] == player_ids[-1] else -rewards[-1]
    return out


def unroll_sequences(actions, rewards, values, pi_probs, unroll_steps: int):
    """pipeline.py:709-767 without the int8 cast of the actions; returns stacked arrays [T, K(, A)]."""
    T, K = len(actions), unroll_steps
    A = len(pi_probs[0])
    a = list(actions) + [0] * K
    r = list(rewards) + [0] * K
    v = list(values) + [0] * K
    p = list(pi_probs) + [np.ones(A, dtype=np.float64) / A] * K
    return (np.array([a[t:t + K] for t in range(T)], dtype=np.int32),
            np.array([r[t:t + K] for t in range(T)], dtype=np.float32),
            np.array([v[t:t + K] for t in range(T)], dtype=np.float32),
            np.array([p[t:t + K] for t in range(T)], dtype=np.float32))
