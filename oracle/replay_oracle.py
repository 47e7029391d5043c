"""TEST INFRASTRUCTURE -- CPU restatement of the reference's replay sampling (muzero/replay.py:38-142).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product
(muzero_b200/replay.py + csrc/replay.cu) never does.

Parity is PINNED: tests/golden/make_golden_replay.py imports the unmodified reference ``PrioritizedReplay`` (with a
stand-in for the absent ``snappy`` module), records indices / weights / RNG states of seeded sampling sessions and
checks this restatement against them (tests/golden/replay_golden.npz, replayed by tests/test_replay.py).

What the reference does, decoded:
* uniform (priority_exponent == 0, the default of every run_training.py): ``self._random_state.uniform(0, size,
  size=batch).astype(np.int64)`` = ``trunc(0.0 + size * u_k)`` with u_k consecutive legacy-MT19937 doubles of the
  replay's OWN RandomState (replay.py:90); weights are float32 ones.
* prioritized: ``priorities[:size] ** exponent`` (float32 power, the python float is weak), ``/ np.sum`` (float32
  pairwise sum, float32 division), then ``np.random.choice(..., p=probs)`` on the GLOBAL numpy stream
  (replay.py:96, not the replay's own state): p converted to float64, ``cdf = p.cumsum(); cdf /= cdf[-1]``,
  ``searchsorted(cdf, u, 'right')`` for ``batch`` consecutive doubles; importance weights
  ``((1.0 / size) / probs[idx]) ** beta`` in float32, divided by their maximum (replay.py:99-100).
* storage is a ring: item k lands in slot ``k % capacity`` (replay.py:76-79).
"""
import numpy as np


def ring_slots(num_added: int, n_new: int, capacity: int) -> np.ndarray:
    """Slots the next ``n_new`` items land in (replay.py:76-79)."""
    return (num_added + np.arange(n_new, dtype=np.int64)) % capacity


def sample_uniform(size: int, batch: int, own_state: np.random.RandomState):
    """replay.py:89-91."""
    u = own_state.random_sample(batch)                     # legacy MT19937 doubles, two 32-bit draws each
    idx = (0.0 + float(size) * u).astype(np.int64)         # RandomState.uniform: low + (high - low) * u, truncated
    return idx, np.ones(batch, dtype=np.float32)


def priority_probs(priorities: np.ndarray, size: int, exponent: float) -> np.ndarray:
    """replay.py:93-94, float32 throughout."""
    assert priorities.dtype == np.float32
    pr = priorities[:size] ** exponent
    return pr / np.sum(pr)


def sample_prioritized(priorities: np.ndarray, size: int, batch: int, exponent: float, beta: float,
                       global_state: np.random.RandomState):
    """replay.py:93-100 with ``np.random.choice`` (legacy ``RandomState.choice`` with p, replace=True) spelled out."""
    probs = priority_probs(priorities, size, exponent)
    cdf = np.cumsum(probs.astype(np.float64))
    cdf /= cdf[-1]
    u = global_state.random_sample(batch)
    idx = cdf.searchsorted(u, side='right').astype(np.int64)
    w = ((1.0 / size) / probs[idx]) ** beta
    w /= np.max(w)
    return idx, w.astype(np.float32)
