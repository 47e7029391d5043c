"""Generate the golden search fixtures by running the UNMODIFIED reference.

Run in the build container (the only place ``/root/reference`` exists):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py [--sweep N]

For every case it runs ``muzero.mcts.uct_search`` from /root/reference with a
deterministic stub network (oracle/stubnet.py:HashStub), records the network
outputs in call order, the returned (action, pi, root value), the whole tree
(N, W, reward, parent, move per expanded node, in expansion order) and the
final state of numpy's global MT19937 stream, and writes
``tests/golden/mcts_golden.npz``.  It also checks the oracle restatement
(oracle/mcts_oracle.py) against the reference bit-for-bit on every case plus
``--sweep`` extra randomized ones, and fails loudly on any mismatch.

The reference is executed with the container's numpy (>= 2, NEP-50 promotion);
that execution — not numpy 1.21 — is the parity target.
"""
from __future__ import annotations

import argparse
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')
sys.dont_write_bytecode = True

from muzero import mcts as ref_mcts                      # noqa: E402  (the reference)
from muzero.config import MuZeroConfig, KnownBounds      # noqa: E402

from oracle import mcts_oracle                           # noqa: E402
from oracle.stubnet import HashStub, ReplayStub          # noqa: E402


def make_config(c):
    cfg = MuZeroConfig(discount=c['discount'], dirichlet_alpha=c['alpha'], num_simulations=c['sims'], batch_size=1,
                       td_steps=0, lr_init=0.0, lr_milestones=[], visit_softmax_temperature_fn=None,
                       known_bounds=KnownBounds(-1, 1) if c['bounds'] else None, is_board_game=c['board'])
    cfg.root_exploration_eps = c['eps']
    return cfg


def run_reference(c, net):
    """Run the reference search, capturing its Node objects in expansion order."""
    order = []
    orig_expand = ref_mcts.Node.expand

    def spy(self, *a, **k):
        order.append(self)
        return orig_expand(self, *a, **k)

    ref_mcts.Node.expand = spy
    try:
        np.random.seed(c['seed'])
        state = np.zeros((2, 2), dtype=np.float32)
        action, pi, rootq = ref_mcts.uct_search(state, net, 'cpu', make_config(c), c['temperature'], c['mask'],
                                                c['players'][0], c['players'][1], c['deterministic'])
    finally:
        ref_mcts.Node.expand = orig_expand
    idx = {id(n): i for i, n in enumerate(order)}
    tree = dict(
        N=np.array([n.N for n in order], np.int32),
        W=np.array([n.W for n in order], np.float64),
        R=np.array([n.reward for n in order], np.float64),
        parent=np.array([idx[id(n.parent)] if n.parent is not None else -1 for n in order], np.int32),
        move=np.array([n.move if n.move is not None else -1 for n in order], np.int32),
        prior=np.array([ch.prior for ch in order[0].children]),
    )
    st = np.random.get_state()
    rng_end = (int(st[2]), zlib.crc32(np.asarray(st[1], np.uint32).tobytes()))
    return int(action), np.asarray(pi), float(rootq), tree, rng_end


def run_oracle(c, net, noise=None):
    np.random.seed(c['seed'])
    state = np.zeros((2, 2), dtype=np.float32)
    a, pi, q, tr = mcts_oracle.uct_search(state, net, 'cpu', make_config(c), c['temperature'], c['mask'],
                                          c['players'][0], c['players'][1], c['deterministic'],
                                          noise=noise, return_trace=True)
    st = np.random.get_state()
    return a, pi, q, tr, (int(st[2]), zlib.crc32(np.asarray(st[1], np.uint32).tobytes()))


def bits(x):
    return np.asarray(x, dtype=np.float64).view(np.uint64)


def compare(c, ref, orc, what):
    a, pi, q, tree, rng_end = ref
    a2, pi2, q2, tr, rng_end2 = orc
    assert a == a2, (what, 'action', a, a2)
    e = max(1.0, min(5.0, 1.0 / c['temperature'])) if c['temperature'] > 0 else 1.0
    if e == int(e):
        assert pi.dtype == pi2.dtype and np.array_equal(bits(pi), bits(pi2)), (what, 'pi', pi, pi2)
    else:   # np.power(float exponent) is SVML-or-libm by CPU: not bit-reproducible in the reference itself
        assert np.allclose(pi, pi2, rtol=1e-15, atol=0, equal_nan=True), (what, 'pi~', pi, pi2)
    assert bits(q) == bits(q2), (what, 'root value', q, q2)
    assert tr.num_nodes == len(tree['N']), (what, 'nodes')
    for k, v in (('N', tr.N), ('parent', tr.parent), ('move', tr.move)):
        assert np.array_equal(tree[k], v), (what, k)
    assert np.array_equal(bits(tree['W']), bits(tr.W)), (what, 'W')
    assert np.array_equal(bits(tree['R']), bits(tr.R)), (what, 'R')
    assert tree['prior'].dtype == tr.prior.dtype, (what, 'prior dtype', tree['prior'].dtype, tr.prior.dtype)
    assert np.array_equal(tree['prior'].view(np.uint8), tr.prior.view(np.uint8)), (what, 'prior')
    assert rng_end == rng_end2, (what, 'rng stream', rng_end, rng_end2)


def random_case(rs, i, small=False, temps=(0.0, 0.1, 0.25, 0.5, 1.0)):
    board = bool(rs.randint(2))
    A = int(rs.choice([2, 4, 10, 18] if small else [2, 3, 10, 18, 82]))
    sims = int(rs.choice([25, 50] if small else [25, 50, 200]))
    deterministic = bool(rs.randint(2))
    style = rs.randint(4)
    if style == 0:
        logits = rs.standard_normal(A) * 2.0
    elif style == 1:
        logits = np.zeros(A)                                   # uniform prior: every U ties
    elif style == 2:
        logits = np.round(rs.standard_normal(A))               # repeated priors: partial ties
    else:
        logits = rs.standard_normal(A) * 6.0                   # peaked
    e = np.exp(logits - logits.max())
    root_pi = (e / e.sum()).astype(np.float32)
    mstyle = rs.randint(4)
    if mstyle == 0:
        mask = np.ones(A, dtype=bool)
    elif mstyle == 3:
        mask = None
    else:
        mask = rs.rand(A) < 0.6
        if not mask.any():
            mask[rs.randint(A)] = True
    players = [(1, 2), (2, 1)][rs.randint(2)] if board else (1, 1)
    if not board and rs.rand() < 0.15:
        players = (1, 2)                                       # exercised by the reference code, keep it honest
    return dict(
        name=f'case{i}', board=board, A=A, sims=sims, deterministic=deterministic,
        discount=1.0 if board else float(rs.choice([0.997, 0.9, 1.0])),
        alpha=float(rs.choice([0.03, 0.25, 1.0])), eps=0.25,
        bounds=bool(board or rs.rand() < 0.2),
        temperature=float(rs.choice(temps)),
        mask=mask, players=players, seed=int(rs.randint(1 << 30)),
        root_pi=root_pi,
        value_scale=float(rs.choice([0.5, 1.0, 3.0])),         # 3.0 blows through known_bounds
        reward_scale=float(rs.choice([0.0, 0.0, 0.7, 2.0])),
        quantise=int(rs.choice([0, 0, 2, 4])),
    )


def run_case(c):
    gen = HashStub(c['root_pi'], c['value_scale'], c['reward_scale'], c['quantise'], seed=c['seed'])
    ref = run_reference(c, gen)
    rewards = np.array([r for r, _ in gen.calls], np.float32)
    values = np.array([v for _, v in gen.calls], np.float32)
    # 1. oracle, own stub run
    gen2 = HashStub(c['root_pi'], c['value_scale'], c['reward_scale'], c['quantise'], seed=c['seed'])
    compare(c, ref, run_oracle(c, gen2), c['name'] + '/hash')
    # 2. oracle fed the RECORDED outputs (what the GPU tests do)
    rep = ReplayStub(c['root_pi'], rewards, values, ref[3]['parent'], ref[3]['move'])
    compare(c, ref, run_oracle(c, rep), c['name'] + '/replay')
    # 3. the reference itself fed the recording must reproduce itself
    rep2 = ReplayStub(c['root_pi'], rewards, values)
    ref2 = run_reference(c, rep2)
    assert ref2[0] == ref[0] and np.array_equal(ref2[3]['N'], ref[3]['N'])
    return ref, rewards, values


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--sweep', type=int, default=150, help='extra randomized reference-vs-oracle cases (not stored)')
    ap.add_argument('--stored', type=int, default=48)
    args = ap.parse_args()

    rs = np.random.RandomState(20261017)
    store = {}
    names = []
    for i in range(args.stored):
        c = random_case(rs, i)
        (a, pi, q, tree, rng_end), rewards, values = run_case(c)
        p = c['name']
        names.append(p)
        meta = np.array([c['board'], c['A'], c['sims'], c['deterministic'], c['bounds'], c['players'][0],
                         c['players'][1], c['seed'], c['mask'] is not None, a, rng_end[0], rng_end[1]], np.int64)
        fl = np.array([c['discount'], c['alpha'], c['eps'], c['temperature'], q], np.float64)
        store.update({
            f'{p}_meta': meta, f'{p}_fl': fl, f'{p}_root_pi': c['root_pi'],
            f'{p}_mask': (c['mask'] if c['mask'] is not None else np.ones(c['A'], bool)),
            f'{p}_rewards': rewards, f'{p}_values': values, f'{p}_pi': pi,
            f'{p}_N': tree['N'], f'{p}_W': tree['W'], f'{p}_R': tree['R'],
            f'{p}_parent': tree['parent'], f'{p}_move': tree['move'], f'{p}_prior': tree['prior'],
        })
        print(f"{p}: board={c['board']} A={c['A']} sims={c['sims']} det={c['deterministic']} T={c['temperature']} "
              f"-> action {a}, rootQ {q:+.6f}, rng pos {rng_end[0]}")
    store['names'] = np.array(names)
    out = os.path.join(HERE, 'mcts_golden.npz')
    np.savez_compressed(out, **store)
    print(f'wrote {out} ({os.path.getsize(out)} bytes, {len(names)} cases)')

    for i in range(args.sweep):
        run_case(random_case(rs, 1000 + i, small=(i % 3 != 0), temps=(0.0, 0.1, 0.25, 0.5, 1.0, 0.3, 0.7)))
    print(f'sweep: {args.sweep} extra randomized cases, oracle == reference bit-for-bit')


if __name__ == '__main__':
    main()
