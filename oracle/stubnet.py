"""TEST INFRASTRUCTURE ONLY — duck-typed stand-ins for ``MuZeroNet``.

``uct_search`` (mcts.py:355-356, 382-384) only needs an object with
``initial_inference(x)`` / ``recurrent_inference(hidden, action)`` returning a
record with ``hidden_state, reward, pi_probs, value`` (network.py:25-30).

* ``HashStub`` invents network outputs as a pure function of the action path,
  using integer arithmetic only, so it behaves identically on every machine.
* ``ReplayStub`` hands back a recorded sequence of (reward, value) in call
  order and checks the (parent path, action) it is asked about — the
  "fed identical network outputs" harness of SURVEY.md §8c (T2).
"""
from __future__ import annotations

from typing import NamedTuple

import numpy as np


class NetworkOutputs(NamedTuple):
    hidden_state: np.ndarray
    reward: float
    pi_probs: np.ndarray
    value: float


def _mix(h: int, a: int) -> int:
    h = (h * 1103515245 + (a + 1) * 2654435761 + 12345) & 0xFFFFFFFF
    h ^= h >> 15
    h = (h * 2246822519) & 0xFFFFFFFF
    h ^= h >> 13
    return h


def _unit(h: int) -> float:
    """Map a 32-bit hash to a float32-representable value in [-1, 1)."""
    return float(np.float32(((h >> 8) / float(1 << 23)) - 1.0))


class HashStub:
    def __init__(self, root_pi: np.ndarray, value_scale: float = 1.0, reward_scale: float = 0.0,
                 quantise: int = 0, seed: int = 0):
        self.root_pi = np.asarray(root_pi, dtype=np.float32)
        self.value_scale, self.reward_scale, self.quantise, self.seed = value_scale, reward_scale, quantise, seed
        self.calls = []          # (reward, value) in call order
        self.root_value = None

    def _q(self, x: float) -> float:
        if self.quantise:       # coarse values make exact Q ties (and RNG tie-breaks) common
            x = round(x * self.quantise) / self.quantise
        return float(np.float32(x))

    def initial_inference(self, x):
        h = _mix(self.seed, 977)
        v = self._q(self.value_scale * _unit(h))
        self.root_value = v
        hid = np.array([h >> 16, h & 0xFFFF], dtype=np.float32)
        return NetworkOutputs(hid, 0.0, self.root_pi.copy(), v)

    def recurrent_inference(self, hidden, action):
        hid = np.asarray(hidden).reshape(-1)
        h = (int(hid[0]) << 16) | int(hid[1])
        a = int(np.asarray(action).reshape(-1)[0])
        h2 = _mix(h, a)
        v = self._q(self.value_scale * _unit(h2))
        r = self._q(self.reward_scale * _unit(_mix(h2, 31337)))
        self.calls.append((r, v))
        return NetworkOutputs(np.array([h2 >> 16, h2 & 0xFFFF], dtype=np.float32), r, self.root_pi.copy(), v)


class ReplayStub:
    """Returns recorded outputs; the hidden state it hands out is the node id,
    so it can assert the search asks about the expected (parent, action)."""

    def __init__(self, root_pi, rewards, values, parents=None, moves=None):
        self.root_pi = np.asarray(root_pi, dtype=np.float32)
        self.rewards, self.values = np.asarray(rewards), np.asarray(values)
        self.parents, self.moves = parents, moves
        self.i = 0

    def initial_inference(self, x):
        self.i = 0
        return NetworkOutputs(np.array([0.0], dtype=np.float32), 0.0, self.root_pi.copy(), 0.0)

    def recurrent_inference(self, hidden, action):
        i = self.i
        self.i += 1
        node = i + 1
        if self.parents is not None:
            par = int(np.asarray(hidden).reshape(-1)[0])
            a = int(np.asarray(action).reshape(-1)[0])
            assert par == int(self.parents[node]) and a == int(self.moves[node]), \
                f'simulation {i}: search expanded ({par},{a}), recording has ({self.parents[node]},{self.moves[node]})'
        return NetworkOutputs(np.array([float(node)], dtype=np.float32), float(self.rewards[i]),
                              self.root_pi.copy(), float(self.values[i]))
