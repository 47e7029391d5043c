"""Record the reference's replay sampling (muzero/replay.py) in seeded sessions.

Run in the build container:  PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_replay.py
``muzero.replay`` imports snappy, which is not installed: a stand-in (compress = tobytes, uncompress = identity) is
injected into sys.modules before the import; the reference file is untouched.  Writes tests/golden/replay_golden.npz and
checks the oracle restatement (oracle/replay_oracle.py) against the recordings on the spot.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')
sys.dont_write_bytecode = True

snappy = types.ModuleType('snappy')
snappy.compress = lambda a: np.ascontiguousarray(a).tobytes()
snappy.uncompress = lambda b: b
sys.modules.setdefault('snappy', snappy)

from muzero.replay import PrioritizedReplay, Transition      # noqa: E402

from oracle import replay_oracle as orc                       # noqa: E402

# (name, capacity, items added, batch, priority exponent, importance exponent, sample calls)
CASES = [
    ('uniform_partial', 1000, 300, 64, 0.0, 0.0, 3),
    ('uniform_wrapped', 500, 1234, 128, 0.0, 0.0, 2),
    ('uniform_large', 100000, 100000, 128, 0.0, 0.0, 2),
    ('prio_exp1', 800, 800, 64, 1.0, 1.0, 2),
    ('prio_exp1_partial', 2000, 777, 128, 1.0, 0.4, 2),
    ('prio_sqrt', 600, 900, 32, 0.5, 0.5, 2),
    ('prio_exp06', 700, 700, 64, 0.6, 0.4, 2),
    ('prio_large', 50000, 50000, 128, 1.0, 1.0, 1),
]


def session(name, capacity, n_items, batch, alpha, beta, calls, seed):
    gen = np.random.RandomState(seed)
    own = np.random.RandomState(seed + 1)
    rep = PrioritizedReplay(capacity, alpha, beta, own)
    prios = (gen.rand(n_items) * 3 + 1e-3).astype(np.float32)
    for k in range(n_items):
        item = Transition(state=np.full((2, 3), k % 127, np.int8), action=np.int32(k), pi_prob=np.float32([k, 1]),
                          value=np.float32(k * 0.5), reward=np.float32(-k))
        rep.add(item, float(prios[k]))
    np.random.seed(seed + 2)                       # the prioritized path draws from the GLOBAL stream
    out = {'priorities': rep._priorities.copy(), 'size': np.int64(rep.size), 'num_added': np.int64(rep.num_added)}
    check_own = np.random.RandomState(seed + 1)
    check_global = np.random.RandomState(seed + 2)
    for c in range(calls):
        batch_t, idx, w = rep.sample(batch)
        out[f'idx{c}'], out[f'w{c}'] = np.asarray(idx, np.int64), np.asarray(w, np.float32)
        out[f'action{c}'] = np.asarray(batch_t.action)
        # the item in slot i is the last k with k % capacity == i
        if alpha == 0:
            oi, ow = orc.sample_uniform(rep.size, batch, check_own)
        else:
            oi, ow = orc.sample_prioritized(rep._priorities, rep.size, batch, alpha, beta, check_global)
        assert np.array_equal(oi, out[f'idx{c}']), (name, c, 'indices')
        assert np.array_equal(ow.view(np.uint32), out[f'w{c}'].view(np.uint32)), (name, c, 'weights')
        if alpha != 0:                               # priorities move between calls like in training
            newp = (gen.rand(batch) * 2 + 1e-3).astype(np.float32)
            rep.update_priorities(idx, newp)
            out[f'newp{c}'] = newp
    out['own_state_key'], out['own_state_pos'] = own.get_state()[1].copy(), np.int64(own.get_state()[2])
    g = np.random.get_state()
    out['global_state_key'], out['global_state_pos'] = g[1].copy(), np.int64(g[2])
    assert np.array_equal(check_own.get_state()[1], out['own_state_key']) and check_own.get_state()[2] == out['own_state_pos']
    assert np.array_equal(check_global.get_state()[1], g[1]) and check_global.get_state()[2] == g[2]
    # ring: slot of item k
    last = {}
    for k in range(n_items):
        last[k % capacity] = k
    assert np.array_equal(orc.ring_slots(0, n_items, capacity), np.arange(n_items) % capacity)
    for c in range(calls):
        want = np.array([last[int(i)] for i in out[f'idx{c}']], np.int32)
        assert np.array_equal(out[f'action{c}'].astype(np.int32), want), (name, 'ring contents')
    return out


def main():
    rec = {}
    meta = []
    for n, case in enumerate(CASES):
        name, capacity, n_items, batch, alpha, beta, calls = case
        out = session(*case, seed=1000 + 17 * n)
        for k, v in out.items():
            rec[f'{name}/{k}'] = v
        meta.append((name, capacity, n_items, batch, alpha, beta, calls, 1000 + 17 * n))
        print('recorded', name)
    rec['cases'] = np.array([repr(m) for m in meta])
    np.savez_compressed(os.path.join(HERE, 'replay_golden.npz'), **rec)
    print('oracle == reference on', len(CASES), 'sessions; wrote replay_golden.npz')


if __name__ == '__main__':
    main()
