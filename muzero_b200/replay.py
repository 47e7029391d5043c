"""Device-resident replay: ``PrioritizedReplay`` of the reference (muzero/replay.py:38-142) with its storage, its
priorities and its sampling on the GPU.

Reference interface mirrored here (same names, argument meaning, error behaviour):
  replay.py:41-66    PrioritizedReplay(capacity, priority_exponent, importance_sampling_exponent, random_state)
  replay.py:68-79    add(item, priority)            -> ring slot num_added % capacity
  replay.py:81-83    get(indices)
  replay.py:85-105   sample(batch_size)             -> (Transition batch, indices, weights)
  replay.py:107-114  update_priorities(indices, priorities)
  replay.py:116-142  num_added / size / capacity / reset / get_state / set_state

What changes: items live in HBM as struct-of-arrays rows (no snappy: the reference compresses states to fit a million
of them in host RAM; 1e6 Gomoku samples are 0.73 GB of int8 here), ``add_batch`` takes the device tensors the
self-play kernels emit (``selfplay.Samples``), ``sample`` returns device tensors that ``training.calc_loss`` consumes
without a host round trip, and the index stream is produced by ``csrc/replay.cu`` from numpy's legacy MT19937 state
kept on the device: the uniform path continues the replay's own ``random_state`` exactly like
``RandomState.uniform`` (replay.py:90), the prioritized path draws from numpy's live GLOBAL stream like the reference's
``np.random.choice`` (replay.py:96) -- loaded to the device and written back around every ``sample`` -- or, when
``global_state`` is given, from that private stream kept on the device.  ``get_state()['random_state']`` returns
the stream so that a host ``RandomState`` can carry on from it.
"""
from __future__ import annotations

from typing import Any, List, Mapping, Optional, Sequence, Text, Tuple

import numpy as np
import torch

from . import _lib
from .training import Transition

_FIELDS = ('state', 'action', 'pi_prob', 'value', 'reward')


class _DeviceStream:
    """numpy legacy MT19937 state on the device (key u32[624], pos i32[1])."""

    def __init__(self, rs: np.random.RandomState, device) -> None:
        self.device = device
        self.key = torch.empty(624, dtype=torch.int32, device=device)
        self.pos = torch.empty(1, dtype=torch.int32, device=device)
        self.set(rs.get_state())

    def set(self, state) -> None:
        assert state[0] == 'MT19937'
        self.key.copy_(torch.from_numpy(np.asarray(state[1], dtype=np.uint32).view(np.int32).copy()))
        self.pos.fill_(int(state[2]))
        self._tail = tuple(state[3:]) if len(state) > 3 else (0, 0.0)

    def get(self):
        key = self.key.cpu().numpy().view(np.uint32).copy()
        return ('MT19937', key, int(self.pos.item())) + tuple(self._tail)


class DeviceReplay:
    """Prioritized replay with circular struct-of-arrays storage in device memory."""

    def __init__(self, capacity: int, priority_exponent: float, importance_sampling_exponent: float,
                 random_state: np.random.RandomState, device='cuda',
                 global_state: Optional[np.random.RandomState] = None) -> None:
        if capacity <= 0:
            raise ValueError(f'Expect capacity to be a positive integer, got {capacity}')      # replay.py:55-56
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('DeviceReplay lives on a CUDA device (no CPU fallback)')
        _lib.lib()                                        # fail loudly if the extension is missing
        self.structure = Transition(None, None, None, None, None)
        self._capacity = int(capacity)
        self._num_added = 0
        self._priority_exponent = float(priority_exponent)
        self._importance_sampling_exponent = float(importance_sampling_exponent)
        with torch.cuda.device(self.device):
            self._priorities = torch.zeros(self._capacity, dtype=torch.float32, device=self.device)
            self._own = _DeviceStream(random_state, self.device)
            # replay.py:96 draws from numpy's LIVE global stream.  Default (global_state=None): every prioritized
            # sample() loads np.random's current state, draws on the device and writes the advanced state back, so
            # the replay and the drop-in uct_search (which consumes the same global stream) never replay each
            # other's numbers.  An explicit `global_state` gives the replay a private device-resident stream instead
            # (no host round trip per sample).
            self._follow_global = global_state is None
            self._global = _DeviceStream(global_state if global_state is not None else _global_numpy_state(),
                                         self.device)
        self._storage: Optional[dict] = None              # field -> tensor [capacity, *item shape]
        self._scratch = None

    # ---- storage ---------------------------------------------------------------------------------------------
    def _ensure_storage(self, fields: dict) -> None:
        if self._storage is not None:
            for k, v in fields.items():
                s = self._storage[k]
                if tuple(s.shape[1:]) != tuple(v.shape[1:]) or s.dtype != v.dtype:
                    raise ValueError(f'replay field {k}: item shape/dtype {tuple(v.shape[1:])}/{v.dtype} differs from '
                                     f'the stored {tuple(s.shape[1:])}/{s.dtype}')
            return
        with torch.cuda.device(self.device):
            self._storage = {k: torch.zeros((self._capacity,) + tuple(v.shape[1:]), dtype=v.dtype, device=self.device)
                             for k, v in fields.items()}

    def add_batch(self, items, priorities) -> None:
        """``add`` for n items at once: ``items`` is a Transition / selfplay.Samples of tensors with a leading n
        (device tensors stay on the device), ``priorities`` n floats.  Item i lands in slot
        ``(num_added + i) % capacity``, exactly where n calls of the reference's ``add`` put it."""
        fields = {k: torch.as_tensor(getattr(items, k)) for k in _FIELDS}
        n = int(fields['state'].shape[0])
        pr = torch.as_tensor(priorities)
        if pr.numel() != n:
            raise ValueError(f'{n} items but {pr.numel()} priorities')
        if n == 0:
            return
        if not bool(torch.isfinite(pr).all()) or bool((pr < 0).any()):
            raise ValueError('Priority must be finite and positive.')                          # replay.py:72-73
        lib, st = _lib.lib(), None
        with torch.cuda.device(self.device):
            fields = {k: v.to(self.device).contiguous() for k, v in fields.items()}
            self._ensure_storage(fields)
            st = _lib.current_stream()
            pr32 = pr.to(device=self.device, dtype=torch.float32).contiguous()
            # more items than slots: only the last `capacity` of them survive (the ring overwrites the rest)
            skip = max(0, n - self._capacity)
            start = (self._num_added + skip) % self._capacity
            m = n - skip
            for k, v in fields.items():
                row = v[0].numel() * v.element_size()
                _lib.check(lib.mz_replay_scatter(v[skip:].contiguous().data_ptr(), self._storage[k].data_ptr(), m, row,
                                                 start, self._capacity, st))
            _lib.check(lib.mz_replay_scatter(pr32[skip:].contiguous().data_ptr(), self._priorities.data_ptr(), m, 4,
                                             start, self._capacity, st))
        self._num_added += n

    def add(self, item, priority: float) -> None:
        """Adds single item to replay (replay.py:68-79)."""
        if not np.isfinite(priority) or priority < 0.0:
            raise ValueError('Priority must be finite and positive.')
        one = Transition(*[torch.as_tensor(np.asarray(getattr(item, k)))[None] for k in _FIELDS])
        self.add_batch(one, [float(priority)])

    def get(self, indices) -> Transition:
        """Items by index, stacked on the batch dimension (the reference returns a list, replay.py:81-83, and
        ``sample`` stacks it, replay.py:102-104)."""
        if self._storage is None:
            raise RuntimeError('Replay is empty')
        lib = _lib.lib()
        with torch.cuda.device(self.device):
            idx = torch.as_tensor(indices, dtype=torch.int64).to(self.device).contiguous()
            n = int(idx.numel())
            out = {}
            for k, s in self._storage.items():
                dst = torch.empty((n,) + tuple(s.shape[1:]), dtype=s.dtype, device=self.device)
                row = s[0].numel() * s.element_size()
                _lib.check(lib.mz_replay_gather(s.data_ptr(), idx.data_ptr(), dst.data_ptr(), n, row,
                                                _lib.current_stream()))
                out[k] = dst
        return Transition(**out)

    # ---- sampling --------------------------------------------------------------------------------------------
    def sample(self, batch_size: int) -> Tuple[Transition, torch.Tensor, torch.Tensor]:
        """Samples batch of items from replay, with replacement (replay.py:85-105).  Returns device tensors:
        (Transition batch, indices int64 [batch], weights float32 [batch])."""
        if self.size < batch_size:
            raise RuntimeError(f'Replay only have {self.size} samples, got sample batch size {batch_size}')
        lib = _lib.lib()
        with torch.cuda.device(self.device):
            st = _lib.current_stream()
            idx = torch.empty(batch_size, dtype=torch.int64, device=self.device)
            w = torch.empty(batch_size, dtype=torch.float32, device=self.device)
            if self._priority_exponent == 0:
                _lib.check(lib.mz_replay_sample_uniform(self.size, batch_size, self._own.key.data_ptr(),
                                                        self._own.pos.data_ptr(), idx.data_ptr(), w.data_ptr(), st))
            else:
                if self._scratch is None:
                    self._scratch = (torch.empty(self._capacity + 1026, dtype=torch.float32, device=self.device),
                                     torch.empty(self._capacity + self._capacity // 2048 + 1, dtype=torch.float64,
                                                 device=self.device))
                if self._follow_global:
                    self._global.set(np.random.get_state())
                _lib.check(lib.mz_replay_sample_prioritized(
                    self.size, batch_size, self._priorities.data_ptr(), self._priority_exponent,
                    self._importance_sampling_exponent, self._global.key.data_ptr(), self._global.pos.data_ptr(),
                    self._scratch[0].data_ptr(), self._scratch[1].data_ptr(), idx.data_ptr(), w.data_ptr(), st))
                if self._follow_global:
                    np.random.set_state(self._global.get())
        return self.get(idx), idx, w

    def update_priorities(self, indices, priorities) -> None:
        """Updates indices with given priorities (replay.py:107-114)."""
        pr = torch.as_tensor(priorities)
        if not bool(torch.isfinite(pr).all()) or bool((pr < 0).any()):
            raise ValueError('Priorities must be finite and positive.')
        with torch.cuda.device(self.device):
            idx = torch.as_tensor(indices, dtype=torch.int64).to(self.device).contiguous()
            pr32 = pr.to(device=self.device, dtype=torch.float32).contiguous()
            n = min(int(idx.numel()), int(pr32.numel()))                                      # zip() semantics
            if n:
                _lib.check(_lib.lib().mz_replay_update_priorities(self._priorities.data_ptr(), idx.data_ptr(),
                                                                  pr32.data_ptr(), n, _lib.current_stream()))

    # ---- bookkeeping -----------------------------------------------------------------------------------------
    @property
    def num_added(self) -> int:
        """Number of items added into replay."""
        return self._num_added

    @property
    def size(self) -> int:
        """Number of items currently contained in replay."""
        return min(self._num_added, self._capacity)

    @property
    def capacity(self) -> int:
        """Total capacity of replay (max number of items stored at any one time)."""
        return self._capacity

    @property
    def priorities(self) -> torch.Tensor:
        return self._priorities

    def reset(self) -> None:
        """Reset the state of replay (replay.py:131-133)."""
        self._num_added = 0

    def get_state(self) -> Mapping[Text, Any]:
        """Replay state as a dictionary of host arrays (replay.py:135-137) plus the two sampling streams."""
        storage = None if self._storage is None else {k: v.cpu().numpy() for k, v in self._storage.items()}
        return {'num_added': self._num_added, 'storage': storage, 'priorities': self._priorities.cpu().numpy(),
                'random_state': self._own.get(), 'global_random_state': self._global.get()}

    def set_state(self, state: Mapping[Text, Any]) -> None:
        """Sets replay state from a (potentially de-serialized) dictionary (replay.py:139-142)."""
        self._num_added = int(state['num_added'])
        with torch.cuda.device(self.device):
            self._priorities.copy_(torch.as_tensor(np.asarray(state['priorities'], dtype=np.float32)))
            if state.get('storage') is not None:
                self._storage = {k: torch.as_tensor(np.asarray(v)).to(self.device).contiguous()
                                 for k, v in state['storage'].items()}
            if 'random_state' in state:
                self._own.set(state['random_state'])
            if 'global_random_state' in state:
                self._global.set(state['global_random_state'])


def _global_numpy_state() -> np.random.RandomState:
    rs = np.random.RandomState()
    rs.set_state(np.random.get_state())
    return rs
