"""Replay sampling: the oracle against the recordings of the reference's PrioritizedReplay
(tests/golden/replay_golden.npz, made by tests/golden/make_golden_replay.py), and -- on the GPU -- DeviceReplay
(csrc/replay.cu behind the C ABI) against the same recordings.  Bar: bit-exact indices, weights and RNG streams for
the uniform path and for priority / importance exponents with an exact numpy path (1, 0.5); powf exponents: weights
to 1e-6 relative, indices equal except where a cdf entry moved by an ulp."""
import ast
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import replay_oracle as orc

G = np.load(os.path.join(ROOT, 'tests', 'golden', 'replay_golden.npz'), allow_pickle=False)
CASES = [ast.literal_eval(str(c)) for c in G['cases']]
IDS = [c[0] for c in CASES]


def _items(n_items, seed):
    gen = np.random.RandomState(seed)
    prios = (gen.rand(n_items) * 3 + 1e-3).astype(np.float32)
    k = np.arange(n_items)
    state = np.broadcast_to((k % 127).astype(np.int8)[:, None, None], (n_items, 2, 3)).copy()
    pi = np.stack([k.astype(np.float32), np.ones(n_items, np.float32)], 1)
    return gen, prios, dict(state=state, action=k.astype(np.int32), pi_prob=pi, value=(k * 0.5).astype(np.float32),
                            reward=(-k).astype(np.float32))


@pytest.mark.parametrize('case', CASES, ids=IDS)
def test_oracle_replays_the_reference_recordings(case):
    name, capacity, n_items, batch, alpha, beta, calls, seed = case
    p = G[f'{name}/priorities'].copy()
    size = int(G[f'{name}/size'])
    own, glob = np.random.RandomState(seed + 1), np.random.RandomState(seed + 2)
    for c in range(calls):
        if alpha == 0:
            idx, w = orc.sample_uniform(size, batch, own)
        else:
            idx, w = orc.sample_prioritized(p, size, batch, alpha, beta, glob)
        assert np.array_equal(idx, G[f'{name}/idx{c}'])
        assert np.array_equal(w.view(np.uint32), G[f'{name}/w{c}'].view(np.uint32))
        if alpha != 0:
            for i, v in zip(idx, G[f'{name}/newp{c}']):
                p[i] = v
    assert np.array_equal(own.get_state()[1], G[f'{name}/own_state_key'])
    assert own.get_state()[2] == int(G[f'{name}/own_state_pos'])
    assert np.array_equal(glob.get_state()[1], G[f'{name}/global_state_key'])
    assert glob.get_state()[2] == int(G[f'{name}/global_state_pos'])


def test_oracle_pairwise_and_ring_helpers():
    assert np.array_equal(orc.ring_slots(7, 5, 10), [7, 8, 9, 0, 1])
    gen = np.random.RandomState(0)
    p = gen.rand(5000).astype(np.float32)
    probs = orc.priority_probs(p, 4000, 1.0)
    assert probs.dtype == np.float32 and abs(float(probs.sum()) - 1.0) < 1e-5


def test_device_replay_argument_errors_need_no_gpu():
    from muzero_b200.replay import DeviceReplay
    with pytest.raises(ValueError):
        DeviceReplay(0, 0.0, 0.0, np.random.RandomState(1))               # replay.py:55-56
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            DeviceReplay(10, 0.0, 0.0, np.random.RandomState(1), device='cpu')


@pytest.mark.gpu
@pytest.mark.parametrize('case', CASES, ids=IDS)
def test_device_replay_matches_the_reference_recordings(case):
    from muzero_b200.replay import DeviceReplay
    from muzero_b200.training import Transition
    name, capacity, n_items, batch, alpha, beta, calls, seed = case
    gen, prios, items = _items(n_items, seed)
    rep = DeviceReplay(capacity, alpha, beta, np.random.RandomState(seed + 1),
                       global_state=np.random.RandomState(seed + 2))
    # ragged chunks (one of them larger than what is left before the ring wraps), the last few one by one
    cuts = sorted(set([0, n_items // 7, n_items // 3, max(0, n_items - 3), n_items]))
    for a, b in zip(cuts[:-1], cuts[1:]):
        if n_items - b < 3 and b - a <= 3:
            for k in range(a, b):
                rep.add(Transition(**{f: v[k] for f, v in items.items()}), float(prios[k]))
        else:
            rep.add_batch(Transition(**{f: torch.from_numpy(v[a:b]).cuda() for f, v in items.items()}), prios[a:b])
    assert rep.num_added == int(G[f'{name}/num_added']) and rep.size == int(G[f'{name}/size'])
    assert np.array_equal(rep.priorities.cpu().numpy(), G[f'{name}/priorities'])
    exact = alpha in (0.0, 1.0, 0.5, 2.0) and (alpha == 0.0 or beta in (1.0, 0.5, 2.0))
    for c in range(calls):
        tr, idx, w = rep.sample(batch)
        idx_h, w_h = idx.cpu().numpy(), w.cpu().numpy()
        want_idx, want_w = G[f'{name}/idx{c}'], G[f'{name}/w{c}']
        if exact:
            assert np.array_equal(idx_h, want_idx)
            assert np.array_equal(w_h.view(np.uint32), want_w.view(np.uint32))
            assert np.array_equal(tr.action.cpu().numpy(), G[f'{name}/action{c}'].astype(np.int32))
            k = tr.action.cpu().numpy().astype(np.int64)
            assert np.array_equal(tr.state.cpu().numpy()[:, 0, 0], (k % 127).astype(np.int8))
            assert np.array_equal(tr.value.cpu().numpy(), (k * 0.5).astype(np.float32))
        else:
            same = idx_h == want_idx
            assert same.mean() >= 0.97
            assert np.allclose(w_h[same], want_w[same], rtol=2e-6, atol=0)
        if alpha != 0:
            rep.update_priorities(torch.from_numpy(want_idx).cuda(), G[f'{name}/newp{c}'])
    st = rep.get_state()
    assert np.array_equal(st['random_state'][1], G[f'{name}/own_state_key'])
    assert st['random_state'][2] == int(G[f'{name}/own_state_pos'])
    assert np.array_equal(st['global_random_state'][1], G[f'{name}/global_state_key'])
    assert st['global_random_state'][2] == int(G[f'{name}/global_state_pos'])


@pytest.mark.gpu
def test_device_replay_state_round_trip_and_errors():
    from muzero_b200.replay import DeviceReplay
    from muzero_b200.training import Transition
    _, prios, items = _items(50, 3)
    rep = DeviceReplay(40, 0.0, 0.0, np.random.RandomState(5))
    with pytest.raises(RuntimeError):
        rep.sample(4)                                                        # replay.py:87-88
    batch = Transition(**{f: torch.from_numpy(v).cuda() for f, v in items.items()})
    with pytest.raises(ValueError):
        rep.add_batch(batch, -prios)                                         # replay.py:72-73
    with pytest.raises(ValueError):
        rep.add(Transition(**{f: v[0] for f, v in items.items()}), float('nan'))
    rep.add_batch(batch, prios)                                              # 50 items into 40 slots
    assert rep.size == 40 and rep.num_added == 50 and rep.capacity == 40
    act = rep.get(np.arange(40)).action.cpu().numpy()
    assert np.array_equal(act, np.where(np.arange(40) < 10, np.arange(40) + 40, np.arange(40)))
    with pytest.raises(ValueError):
        rep.update_priorities([0, 1], [1.0, float('inf')])                   # replay.py:110-111
    a = rep.sample(16)
    state = rep.get_state()
    b = rep.sample(16)
    rep2 = DeviceReplay(40, 0.0, 0.0, np.random.RandomState(99))
    rep2.set_state(state)
    b2 = rep2.sample(16)
    assert np.array_equal(b[1].cpu().numpy(), b2[1].cpu().numpy())
    assert np.array_equal(b[0].state.cpu().numpy(), b2[0].state.cpu().numpy())
    # the host stream continues where the device left off (same draws as numpy from that state)
    host = np.random.RandomState()
    host.set_state(state['random_state'])
    want, _ = orc.sample_uniform(40, 16, host)
    assert np.array_equal(b[1].cpu().numpy(), want) and not np.array_equal(a[1].cpu().numpy(), want)
    rep.reset()
    assert rep.size == 0


@pytest.mark.gpu
def test_self_play_samples_flow_into_replay_and_the_learner():
    """The f-rows together: board self-play kernels -> DeviceReplay -> calc_loss, all on the device."""
    import muzero_b200 as mz
    from muzero_b200.replay import DeviceReplay
    from muzero_b200.selfplay import BatchedBoardEnv, BoardSelfPlay
    from muzero_b200.training import DataParallelLearner
    torch.manual_seed(0)
    net = mz.MuZeroMLPNet((9, 3, 3), 10, 64, 1, 1, 64).cuda().eval()
    cfg = mz.config.make_tictactoe_config(num_training_steps=10, batch_size=32)
    cfg.num_simulations = 8
    env = BatchedBoardEnv(64, 3, 3, 4)
    sp = BoardSelfPlay(net, cfg, env, seed=7)
    rep = DeviceReplay(500, 0.0, 0.0, np.random.RandomState(1))
    for _ in range(12):
        s = sp.play_move()
        if s is not None:
            rep.add_batch(s, s.priority)
    assert rep.size > 32
    tr, idx, w = rep.sample(32)
    assert tr.state.is_cuda and tr.pi_prob.shape == (32, cfg.unroll_steps, 10)
    learner = DataParallelLearner(net, cfg, 'cuda')
    loss, prio = learner.step(tr, w)
    assert np.isfinite(loss) and prio.shape == (32,)
    rep.update_priorities(idx, prio)
    net.eval()
    sp.play_move()                      # the actor searches with the refreshed weights (engine rebuilt)


@pytest.mark.gpu
def test_prioritized_sampling_follows_numpys_live_global_stream():
    """replay.py:96 calls the GLOBAL np.random.choice at every sample(): with no `global_state` the device replay must
    consume np.random's live stream (and leave it advanced), interleaved with other users of that stream."""
    from muzero_b200.replay import DeviceReplay
    from muzero_b200.training import Transition
    _, prios, items = _items(200, 11)
    rep = DeviceReplay(200, 1.0, 1.0, np.random.RandomState(4))
    rep.add_batch(Transition(**{f: torch.from_numpy(v).cuda() for f, v in items.items()}), prios)
    np.random.seed(321)
    twin = np.random.RandomState(321)
    for _ in range(3):
        np.random.random_sample(5); twin.random_sample(5)                    # somebody else draws in between
        _, idx, w = rep.sample(16)
        want_idx, want_w = orc.sample_prioritized(prios.astype(np.float32), 200, 16, 1.0, 1.0, twin)
        assert np.array_equal(idx.cpu().numpy(), want_idx)
        assert np.array_equal(w.cpu().numpy().view(np.uint32), want_w.view(np.uint32))
    a, b = np.random.get_state(), twin.get_state()
    assert np.array_equal(a[1], b[1]) and a[2] == b[2]


@pytest.mark.gpu
@pytest.mark.parametrize('n,tiny', [(100, False), (5000, False), (200_000, False), (3000, True)])
def test_parallel_total_and_scan_are_bit_exact(n, tiny):
    """The prioritized path evaluates numpy's float32 pairwise total in parallel along numpy's own recursion tree and the
    float64 running sum by a parallel scan when no addition can round (else sequentially: `tiny` plants probabilities
    below 2^-29).  Indices and weights must equal the oracle's (numpy's) bit for bit at sizes that exercise several
    levels of the tree / several scan blocks."""
    from muzero_b200.replay import DeviceReplay
    from muzero_b200.training import Transition
    gen = np.random.RandomState(n)
    prios = (gen.rand(n) * 3 + 1e-3).astype(np.float32)
    if tiny:
        prios[::7] = 1e-12
    k = np.arange(n)
    items = dict(state=(k % 127).astype(np.int8)[:, None], action=k.astype(np.int32), pi_prob=k.astype(np.float32)[:, None],
                 value=k.astype(np.float32), reward=k.astype(np.float32))
    rep = DeviceReplay(n, 1.0, 1.0, np.random.RandomState(1), global_state=np.random.RandomState(2))
    rep.add_batch(Transition(**{f: torch.from_numpy(v).cuda() for f, v in items.items()}), prios)
    twin = np.random.RandomState(2)
    for _ in range(3):
        _, idx, w = rep.sample(64)
        want_idx, want_w = orc.sample_prioritized(prios, n, 64, 1.0, 1.0, twin)
        assert np.array_equal(idx.cpu().numpy(), want_idx)
        assert np.array_equal(w.cpu().numpy().view(np.uint32), want_w.view(np.uint32))
