"""Callers of the search path: board environments and trajectory -> targets.

CPU part: the oracle restatement (oracle/selfplay_oracle.py) against recordings of the UNMODIFIED reference
(tests/golden/selfplay_golden.npz, written by tests/golden/make_golden_selfplay.py) and the reference's own
known-answer tests (tests/pipeline_test.py:24-53).  GPU part: the CUDA kernels, through the C ABI, bit-for-bit
against the recordings and against the oracle on random batches.
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, bits
from oracle import selfplay_oracle as orc

Z = np.load(os.path.join(GOLDEN, 'selfplay_golden.npz'))
ENV_NAMES = [str(n) for n in Z['env_names']]
TRAJ_NAMES = [str(n) for n in Z['traj_names']]


def env_case(name):
    return {k: Z[f'env_{name}_{k}'] for k in ('actions', 'obs', 'reward', 'done', 'player', 'mask', 'shape')}


def traj_case(name):
    d = {k: Z[f'{name}_{k}'] for k in ('meta', 'discount', 'rewards', 'roots', 'players', 'actions', 'pis', 'nstep',
                                        'mc', 'prio', 'sa', 'sr', 'sv', 'sp')}
    d['T'], d['A'], d['K'], d['n'], d['board'] = (int(x) for x in d['meta'])
    return d


# --------------------------------------------------------------------------- CPU: oracle vs reference recordings
@pytest.mark.parametrize('name', ENV_NAMES)
def test_env_oracle_replays_reference_games(name):
    c = env_case(name)
    n, k, stack = (int(x) for x in c['shape'])
    o = orc.BoardEnvOracle(n, k, stack)
    assert np.array_equal(o.reset(), c['obs'][0]) and np.array_equal(o.mask, c['mask'][0])
    for t, a in enumerate(c['actions']):
        ob, r, d = o.step(int(a))
        assert np.array_equal(ob, c['obs'][t + 1]) and r == c['reward'][t] and d == c['done'][t]
        assert o.player == c['player'][t + 1] and np.array_equal(o.mask, c['mask'][t + 1])
    with pytest.raises(RuntimeError):
        o.step(int(np.flatnonzero(o.mask)[0]) if o.mask.any() else 0)


@pytest.mark.parametrize('name', TRAJ_NAMES)
def test_target_oracle_matches_reference_recordings(name):
    c = traj_case(name)
    nstep = orc.n_step_target(list(c['rewards']), list(c['roots']), c['n'], float(c['discount'][0]))
    assert np.array_equal(bits(nstep), bits(c['nstep']))
    mc = orc.mc_return_target(list(c['rewards']), list(c['players']))
    assert np.array_equal(bits(mc), bits(c['mc']))
    target = mc if c['board'] else nstep
    sa, sr, sv, sp = orc.unroll_sequences(list(c['actions']), list(c['rewards']), target, list(c['pis']), c['K'])
    assert np.array_equal(sa, c['sa']) and np.array_equal(sr, c['sr']) and np.array_equal(sv, c['sv'])
    assert np.array_equal(sp, c['sp'])


def test_reference_known_answers_for_n_step_target():
    """tests/pipeline_test.py:24-53 of the reference."""
    t = orc.n_step_target([1.0] * 5, [0] * 5, 5, 0.997)
    np.testing.assert_almost_equal(np.array(t), np.array([4.97, 3.982, 2.991, 1.997, 1.0]), decimal=3)
    roots = [0.1 * (i + 1) for i in range(10)]
    t = orc.n_step_target([1.0] * 10, roots, 5, 0.997)
    want = [4.97 + 0.997 ** 5 * roots[i + 5] for i in range(5)] + [4.97, 3.982, 2.991, 1.997, 1.0]
    np.testing.assert_almost_equal(np.array(t), np.array(want), decimal=3)
    with pytest.raises(ValueError):
        orc.n_step_target([1.0], [0.0, 1.0], 5, 0.997)


# --------------------------------------------------------------------------- GPU: kernels through the C ABI
@pytest.mark.gpu
def test_env_kernels_replay_reference_games_in_one_batch():
    """All recorded games of one shape run as ONE batch (ragged lengths: finished games idle)."""
    import muzero_b200 as mz
    for prefix in ('tictactoe', 'gomoku9', 'gomoku5'):
        cases = [env_case(n) for n in ENV_NAMES if n.startswith(prefix + '_')]
        n, k, stack = (int(x) for x in cases[0]['shape'])
        G = len(cases)
        env = mz.BatchedBoardEnv(G, n, k, stack)
        for g, c in enumerate(cases):
            assert np.array_equal(env.obs[g].cpu().numpy(), c['obs'][0].astype(np.float32))
        longest = max(len(c['actions']) for c in cases)
        for t in range(longest):
            act = np.array([c['actions'][t] if t < len(c['actions']) else 0 for c in cases], np.int32)
            obs, reward, done, mover = env.step(torch.from_numpy(act).cuda())
            obs, reward, done, mover = obs.cpu().numpy(), reward.cpu().numpy(), done.cpu().numpy(), mover.cpu().numpy()
            mask, player = env.actions_mask.cpu().numpy(), env.current_player.cpu().numpy()
            for g, c in enumerate(cases):
                if t < len(c['actions']):
                    assert np.array_equal(obs[g], c['obs'][t + 1].astype(np.float32)), (prefix, g, t)
                    assert reward[g] == c['reward'][t] and bool(done[g]) == bool(c['done'][t])
                    assert mover[g] == c['player'][t] and player[g] == c['player'][t + 1]
                    assert np.array_equal(mask[g].astype(bool), c['mask'][t + 1])
                else:
                    assert done[g] == 1 and reward[g] == 0.0
        env.check_errors()
        # an action that was already taken is reported (the reference raises ValueError, env.py:121-122)
        env.reset()
        a = torch.zeros(G, dtype=torch.int32).cuda()
        env.step(a)
        env.step(a)
        with pytest.raises(ValueError):
            env.check_errors()
        # partial reset
        which = torch.zeros(G, dtype=torch.uint8)
        which[0] = 1
        env.reset(which.cuda())
        assert int(env.steps[0]) == 0 and int(env.steps[1]) == 1 and float(env.obs[0, :-1].abs().sum()) == 0.0


@pytest.mark.gpu
def test_target_kernels_match_reference_recordings():
    import muzero_b200 as mz
    cases = [traj_case(n) for n in TRAJ_NAMES]
    for c in cases:
        T, A, K = c['T'], c['A'], c['K']
        Tmax = T + 3
        pad = lambda x, dt: torch.from_numpy(np.concatenate([np.asarray(x), np.zeros(Tmax - T)]).astype(dt)[None]).cuda()
        lengths = torch.tensor([T], dtype=torch.int32).cuda()
        rewards, roots = pad(c['rewards'], np.float64), pad(c['roots'], np.float64)
        players, actions = pad(c['players'], np.int32), pad(c['actions'], np.int32)
        pis = torch.zeros((1, Tmax, A), dtype=torch.float32)
        pis[0, :T] = torch.from_numpy(c['pis'].astype(np.float32))
        pis = pis.cuda()
        nstep, prio_n = mz.n_step_targets(lengths, rewards, roots, c['n'], float(c['discount'][0]))
        assert np.array_equal(bits(nstep[0, :T].cpu().numpy()), bits(c['nstep']))
        mc, prio_m = mz.mc_return_targets(lengths, rewards, players, roots)
        assert np.array_equal(bits(mc[0, :T].cpu().numpy()), bits(c['mc']))
        target, prio = (mc, prio_m) if c['board'] else (nstep, prio_n)
        assert np.array_equal(bits(prio[0, :T].cpu().numpy()), bits(c['prio']))
        assert float(target[0, T:].abs().sum()) == 0.0
        sa, sr, sv, sp, valid = mz.unroll_sequences(lengths, actions, rewards, target, pis, K)
        assert valid[0].cpu().numpy().tolist() == [True] * T + [False] * (Tmax - T)
        assert np.array_equal(sa[0, :T].cpu().numpy(), c['sa']) and np.array_equal(sr[0, :T].cpu().numpy(), c['sr'])
        assert np.array_equal(sv[0, :T].cpu().numpy(), c['sv']) and np.array_equal(sp[0, :T].cpu().numpy(), c['sp'])


@pytest.mark.gpu
def test_target_kernels_vs_oracle_on_a_ragged_batch():
    import muzero_b200 as mz
    gen = np.random.RandomState(7)
    G, Tmax, A, K, n, discount = 300, 50, 10, 5, 5, 0.997
    lens = gen.randint(1, Tmax + 1, size=G).astype(np.int32)
    rewards = np.zeros((G, Tmax)); roots = np.zeros((G, Tmax)); players = np.zeros((G, Tmax), np.int32)
    actions = np.zeros((G, Tmax), np.int32); pis = np.zeros((G, Tmax, A), np.float32)
    for g in range(G):
        T = lens[g]
        rewards[g, :T] = gen.standard_normal(T); roots[g, :T] = gen.standard_normal(T)
        players[g, :T] = 1 + (np.arange(T) % 2); actions[g, :T] = gen.randint(0, A, size=T)
        pis[g, :T] = gen.dirichlet(np.ones(A), size=T)
    d = lambda x: torch.from_numpy(x).cuda()
    nstep, prio = mz.n_step_targets(d(lens), d(rewards), d(roots), n, discount)
    mc, _ = mz.mc_return_targets(d(lens), d(rewards), d(players), d(roots))
    sa, sr, sv, sp, valid = mz.unroll_sequences(d(lens), d(actions), d(rewards), nstep, d(pis), K)
    nstep, prio, mc = nstep.cpu().numpy(), prio.cpu().numpy(), mc.cpu().numpy()
    sa, sr, sv, sp = sa.cpu().numpy(), sr.cpu().numpy(), sv.cpu().numpy(), sp.cpu().numpy()
    for g in range(0, G, 7):
        T = lens[g]
        want = orc.n_step_target(list(rewards[g, :T]), list(roots[g, :T]), n, discount)
        assert np.array_equal(bits(nstep[g, :T]), bits(want))
        assert np.array_equal(bits(prio[g, :T]), bits(np.abs(roots[g, :T] - np.array(want))))
        assert np.array_equal(bits(mc[g, :T]), bits(orc.mc_return_target(list(rewards[g, :T]), list(players[g, :T]))))
        oa, orw, ov, op = orc.unroll_sequences(list(actions[g, :T]), list(rewards[g, :T]), want,
                                               [p.astype(np.float64) for p in pis[g, :T]], K)
        assert np.array_equal(sa[g, :T], oa) and np.array_equal(sr[g, :T], orw) and np.array_equal(sv[g, :T], ov)
        assert np.array_equal(sp[g, :T], op)


@pytest.mark.gpu
def test_board_self_play_loop_on_device():
    """run_self_play for 64 concurrent Tic-Tac-Toe games with the reference checkpoint: the moves the device loop
    played, replayed through the oracle environment, give the same observations / rewards / ends, and the emitted
    samples are the oracle's MC-return targets and unroll windows of those trajectories."""
    import muzero_b200 as mz
    sd = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLDEN, 'ckpt_tictactoe.npz')).items()}
    net = mz.MuZeroMLPNet((9, 3, 3), 10, 256, 1, 1, 64)
    net.load_state_dict(sd)
    net = net.cuda().eval()
    cfg = mz.make_tictactoe_config(use_tensorboard=False)
    G = 64
    env = mz.BatchedBoardEnv(G, 3, 3, 4)
    loop = mz.BoardSelfPlay(net, cfg, env, seed=123)
    oracles = [orc.BoardEnvOracle(3, 3, 4) for _ in range(G)]
    traj = [dict(obs=[], a=[], r=[], root=[], pi=[], pl=[]) for _ in range(G)]
    finished = []
    for move in range(12):
        obs_before = env.obs.cpu().numpy().copy()
        steps_before = loop.steps.copy()
        out = loop.play_move()
        act = loop.t_action.cpu().numpy()
        roots, pis = loop.t_root.cpu().numpy(), loop.t_pi.cpu().numpy()
        ended = []
        for g in range(G):
            t = steps_before[g]
            o = oracles[g]
            assert np.array_equal(obs_before[g], o.observation().astype(np.float32))
            a = int(act[g, t])
            assert o.mask[a], 'the search picked an illegal move'
            tr = traj[g]
            tr['obs'].append(o.observation()); tr['a'].append(a); tr['pl'].append(o.player)
            tr['root'].append(float(roots[g, t])); tr['pi'].append(pis[g, t].astype(np.float64))
            _, r, d = o.step(a)
            tr['r'].append(r)
            if d:
                ended.append(g)
        if not ended:
            assert out is None
            continue
        n = sum(len(traj[g]['a']) for g in ended)
        assert out.state.shape[0] == n and out.priority.shape[0] == n
        k = 0
        for g in ended:                                   # samples come game by game, step by step
            tr = traj[g]
            target = orc.mc_return_target(tr['r'], tr['pl'])
            oa, orw, ov, op = orc.unroll_sequences(tr['a'], tr['r'], target, tr['pi'], cfg.unroll_steps)
            T = len(tr['a'])
            assert np.array_equal(out.state[k:k + T].cpu().numpy(), np.array(tr['obs'], np.int8))
            assert np.array_equal(out.action[k:k + T].cpu().numpy(), oa)
            assert np.array_equal(out.reward[k:k + T].cpu().numpy(), orw)
            assert np.array_equal(out.value[k:k + T].cpu().numpy(), ov)
            assert np.array_equal(out.pi_prob[k:k + T].cpu().numpy(), op)
            assert np.array_equal(bits(out.priority[k:k + T].cpu().numpy()),
                                  bits(np.abs(np.array(tr['root']) - np.array(target))))
            k += T
            finished.append(g)
            oracles[g] = orc.BoardEnvOracle(3, 3, 4)
            traj[g] = dict(obs=[], a=[], r=[], root=[], pi=[], pl=[])
    env.check_errors()
    assert len(finished) >= G and loop.games_finished == len(finished)
