import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu')


class GoldenCase:
    """One recorded run of the reference ``uct_search`` (tests/golden/make_golden.py)."""

    def __init__(self, z, name):
        g = lambda k: z[f'{name}_{k}']
        m = g('meta')
        self.name = name
        self.board, self.A, self.sims = bool(m[0]), int(m[1]), int(m[2])
        self.deterministic, self.bounds = bool(m[3]), bool(m[4])
        self.players = (int(m[5]), int(m[6]))
        self.seed, self.has_mask, self.action = int(m[7]), bool(m[8]), int(m[9])
        self.rng_end = (int(m[10]), int(m[11]))
        fl = g('fl')
        self.discount, self.alpha, self.eps, self.temperature, self.root_value = (float(x) for x in fl)
        self.root_pi, self.mask = g('root_pi'), (g('mask') if self.has_mask else None)
        self.rewards, self.values, self.pi = g('rewards'), g('values'), g('pi')
        self.N, self.W, self.R, self.parent, self.move, self.prior = (g(k) for k in 'N W R parent move prior'.split())

    def config(self):
        from muzero_b200.config import KnownBounds, MuZeroConfig
        cfg = MuZeroConfig(discount=self.discount, dirichlet_alpha=self.alpha, num_simulations=self.sims,
                           batch_size=1, td_steps=0, lr_init=0.0, lr_milestones=[],
                           visit_softmax_temperature_fn=None,
                           known_bounds=KnownBounds(-1, 1) if self.bounds else None, is_board_game=self.board)
        cfg.root_exploration_eps = self.eps
        return cfg


def load_golden_cases():
    z = np.load(os.path.join(GOLDEN, 'mcts_golden.npz'))
    return [GoldenCase(z, str(n)) for n in z['names']]


@pytest.fixture(scope='session')
def golden_cases():
    return load_golden_cases()


def bits(x):
    return np.atleast_1d(np.asarray(x, dtype=np.float64)).view(np.uint64)
