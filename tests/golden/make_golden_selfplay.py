"""Record the reference's board environments and trajectory->target functions on random inputs.

Run in the build container:  PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_selfplay.py
`muzero.games.*` and `muzero.pipeline` import gym / snappy, which are not installed: test-side stand-ins are injected
into sys.modules before the import (the reference files are untouched).  Writes tests/golden/selfplay_golden.npz and
checks the oracle restatement (oracle/selfplay_oracle.py) against the recordings on the spot.
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


class _Env:
    def reset(self, **kwargs):
        return None

    def close(self):
        return None


class _Box:
    def __init__(self, low=None, high=None, shape=None, dtype=None):
        self.low, self.high, self.shape, self.dtype = low, high, shape, dtype


class _Discrete:
    def __init__(self, n):
        self.n = n


gym = types.ModuleType('gym')
gym.Env = _Env
gym.Wrapper = gym.ObservationWrapper = gym.RewardWrapper = object
gym.spaces = types.ModuleType('gym.spaces')
gym.spaces.Box, gym.spaces.Discrete = _Box, _Discrete
gym.utils = types.ModuleType('gym.utils')
gym.utils.seeding = types.ModuleType('gym.utils.seeding')
for name, mod in (('gym', gym), ('gym.spaces', gym.spaces), ('gym.utils', gym.utils),
                  ('gym.utils.seeding', gym.utils.seeding), ('snappy', types.ModuleType('snappy'))):
    sys.modules.setdefault(name, mod)
if not hasattr(np, 'bool8'):
    np.bool8 = np.bool_

from muzero.games.gomoku import GomokuEnv                     # noqa: E402
from muzero.games.tictactoe import TicTacToeEnv               # noqa: E402
from muzero import pipeline as ref_pipe                       # noqa: E402

from oracle import selfplay_oracle as orc                     # noqa: E402

ENVS = {
    'tictactoe': (lambda: TicTacToeEnv(), (3, 3, 4)),
    'gomoku9': (lambda: GomokuEnv(board_size=9, num_to_win=5, stack_history=4), (9, 5, 4)),
    'gomoku5': (lambda: GomokuEnv(board_size=5, num_to_win=4, stack_history=2), (5, 4, 2)),
}


def play(env, gen, resign_prob):
    obs0 = env.reset()
    acts, obs, rew, done, player, mask = [], [obs0], [], [], [env.current_player], [env.actions_mask.copy()]
    while True:
        legal = np.flatnonzero(env.actions_mask[:-1])
        a = int(env.resign_action) if (gen.rand() < resign_prob or len(legal) == 0) else int(gen.choice(legal))
        o, r, d, _ = env.step(a)
        acts.append(a); obs.append(o); rew.append(r); done.append(d)
        player.append(env.current_player); mask.append(env.actions_mask.copy())
        if d:
            break
    return (np.array(acts, np.int32), np.array(obs, np.int8), np.array(rew, np.float64), np.array(done, bool),
            np.array(player, np.int32), np.array(mask, bool))


def main():
    gen = np.random.RandomState(2024)
    out = {}
    names = []
    for ename, (make, (n, k, stack)) in ENVS.items():
        for j in range(12):
            env = make()
            rec = play(env, gen, resign_prob=0.0 if j % 3 else 0.03)
            name = f'{ename}_{j}'
            names.append(name)
            for key, val in zip(('actions', 'obs', 'reward', 'done', 'player', 'mask'), rec):
                out[f'env_{name}_{key}'] = val
            out[f'env_{name}_shape'] = np.array([n, k, stack], np.int32)
            # the oracle restatement against the recording
            o = orc.BoardEnvOracle(n, k, stack)
            assert np.array_equal(o.reset(), rec[1][0])
            for t, a in enumerate(rec[0]):
                ob, r, d = o.step(int(a))
                assert np.array_equal(ob, rec[1][t + 1]) and r == rec[2][t] and d == rec[3][t], (name, t)
                assert o.player == rec[4][t + 1] and np.array_equal(o.mask, rec[5][t + 1])
    out['env_names'] = np.array(names)

    tnames = []
    for j in range(16):
        T = int(gen.randint(1, 40))
        A = int(gen.choice([2, 10, 82]))
        K = int(gen.choice([1, 5]))
        n = int(gen.choice([0, 1, 5, 10]))
        discount = float(gen.choice([1.0, 0.997, 0.9]))
        rewards = [float(x) for x in np.round(gen.standard_normal(T), 3)]
        if j % 2:
            rewards = [0.0] * (T - 1) + [float(gen.choice([-1.0, 0.0, 1.0]))]
        roots = [float(x) for x in gen.standard_normal(T)]
        players = [int(1 + (t % 2)) for t in range(T)]
        actions = [int(x) for x in gen.randint(0, A, size=T)]
        pis = [gen.dirichlet(np.ones(A)) for _ in range(T)]
        nstep = ref_pipe.compute_n_step_target(list(rewards), list(roots), n, discount)
        mc = ref_pipe.compute_mc_return_target(list(rewards), list(players))
        target = mc if j % 2 else nstep
        prio = np.abs(np.array(roots) - np.array(target))
        seqs = list(ref_pipe.make_unroll_sequence([np.zeros(1)] * T, list(actions), list(rewards), list(pis),
                                                  list(target), list(prio), K))
        name = f'traj_{j}'
        tnames.append(name)
        out[f'{name}_meta'] = np.array([T, A, K, n, j % 2], np.int32)
        out[f'{name}_discount'] = np.array([discount])
        out[f'{name}_rewards'] = np.array(rewards); out[f'{name}_roots'] = np.array(roots)
        out[f'{name}_players'] = np.array(players, np.int32); out[f'{name}_actions'] = np.array(actions, np.int32)
        out[f'{name}_pis'] = np.array(pis)
        out[f'{name}_nstep'] = np.array(nstep, np.float64); out[f'{name}_mc'] = np.array(mc, np.float64)
        out[f'{name}_prio'] = prio
        out[f'{name}_sa'] = np.array([s[0].action for s in seqs]).astype(np.int32)     # int8 in the reference (A <= 127 here)
        out[f'{name}_sr'] = np.array([s[0].reward for s in seqs]); out[f'{name}_sv'] = np.array([s[0].value for s in seqs])
        out[f'{name}_sp'] = np.array([s[0].pi_prob for s in seqs])
        # oracle against the recording
        assert orc.n_step_target(rewards, roots, n, discount) == nstep
        assert orc.mc_return_target(rewards, players) == mc
        sa, sr, sv, sp = orc.unroll_sequences(actions, rewards, target, pis, K)
        assert np.array_equal(sa, out[f'{name}_sa']) and np.array_equal(sr, out[f'{name}_sr'])
        assert np.array_equal(sv, out[f'{name}_sv']) and np.array_equal(sp, out[f'{name}_sp'])
    out['traj_names'] = np.array(tnames)
    # the reference's own known-answer tests (tests/pipeline_test.py:24-53)
    t = ref_pipe.compute_n_step_target([1.0] * 5, [0] * 5, 5, 0.997)
    np.testing.assert_almost_equal(np.array(t), np.array([4.97, 3.982, 2.991, 1.997, 1.0]), decimal=3)
    np.savez_compressed(os.path.join(HERE, 'selfplay_golden.npz'), **out)
    print('wrote selfplay_golden.npz:', len(names), 'games,', len(tnames), 'trajectories')


if __name__ == '__main__':
    main()
