"""Search / training hyper-parameters, attribute-compatible with the reference.

Mirror of the *interface* of ``muzero/config.py`` (MuZeroConfig at
config.py:22-103, the four factories at config.py:106-233 and the temperature
schedules at config.py:236-267): ``uct_search`` accepts either this class or
the reference's own ``MuZeroConfig`` instance — it only reads attributes.
"""
from __future__ import annotations

from collections import namedtuple
from typing import Callable, Optional, Sequence

KnownBounds = namedtuple('KnownBounds', ['min', 'max'])


class MuZeroConfig:
    # attribute -> default; anything here can be overridden by keyword
    _DEFAULTS = dict(
        known_bounds=None, num_training_steps=int(1000e3), checkpoint_interval=int(1e3), num_planes=256,
        num_res_blocks=16, hidden_dim=64, value_support_size=1, reward_support_size=1, train_delay=0.0,
        min_replay_size=int(2e4), acc_seq_length=200, clip_grad=False, use_tensorboard=False, is_board_game=False,
    )

    def __init__(self, discount: float, dirichlet_alpha: float, num_simulations: int, batch_size: int, td_steps: int,
                 lr_init: float, lr_milestones: Sequence[int],
                 visit_softmax_temperature_fn: Optional[Callable[[int, int], float]], **overrides) -> None:
        unknown = set(overrides) - set(self._DEFAULTS)
        if unknown:
            raise TypeError(f'unexpected MuZeroConfig arguments: {sorted(unknown)}')
        for k, v in {**self._DEFAULTS, **overrides}.items():
            setattr(self, k, v)
        # self-play / search (read by uct_search: mcts.py:349-369,147-155,193)
        self.visit_softmax_temperature_fn = visit_softmax_temperature_fn
        self.num_simulations = num_simulations
        self.discount = discount
        self.root_dirichlet_alpha = dirichlet_alpha
        self.root_exploration_eps = 0.25
        self.pb_c_base = 19652
        self.pb_c_init = 1.25
        # training
        self.batch_size = batch_size
        self.unroll_steps = 5
        self.td_steps = td_steps
        self.weight_decay = 1e-4
        self.momentum = 0.9
        self.max_grad_norm = 40.0
        self.lr_init = lr_init
        self.lr_decay_rate = 0.1
        self.lr_milestones = lr_milestones


def _stepped(limit_attr: str, limit, late: float):
    def fn(env_steps, training_steps):
        x = env_steps if limit_attr == 'env' else training_steps
        return 1.0 if x < limit else late
    return fn


tictactoe_visit_softmax_temperature_fn = _stepped('env', 6, 0.1)      # config.py:236-241
gomoku_visit_softmax_temperature_fn = _stepped('env', 30, 0.1)        # config.py:244-249


def classic_visit_softmax_temperature_fn(env_steps, training_steps):  # config.py:252-258
    return 1.0 if training_steps < 30000 else (0.5 if training_steps < 60000 else 0.25)


def atari_visit_softmax_temperature_fn(env_steps, training_steps):    # config.py:261-267
    return 1.0 if training_steps < 500e3 else (0.5 if training_steps < 1000e3 else 0.25)


def make_tictactoe_config(num_training_steps=100000, batch_size=128, min_replay_size=10000, use_mlp_net=True,
                          use_tensorboard=True, clip_grad=False) -> MuZeroConfig:
    """config.py:106-136."""
    return MuZeroConfig(
        discount=1.0, dirichlet_alpha=0.25, num_simulations=25, batch_size=batch_size, td_steps=0, lr_init=0.002,
        lr_milestones=[20000], visit_softmax_temperature_fn=tictactoe_visit_softmax_temperature_fn,
        known_bounds=KnownBounds(-1, 1), num_training_steps=num_training_steps,
        num_planes=256 if use_mlp_net else 16, num_res_blocks=0 if use_mlp_net else 2,
        hidden_dim=64 if use_mlp_net else 0, min_replay_size=min_replay_size, checkpoint_interval=500,
        acc_seq_length=9999, clip_grad=clip_grad, use_tensorboard=use_tensorboard, is_board_game=True)


def make_gomoku_config(num_training_steps=1000000, batch_size=128, min_replay_size=10000, use_tensorboard=True,
                       clip_grad=False) -> MuZeroConfig:
    """config.py:139-167."""
    return MuZeroConfig(
        discount=1.0, dirichlet_alpha=0.03, num_simulations=200, batch_size=batch_size, td_steps=0, lr_init=0.002,
        lr_milestones=[200e3, 400e3], visit_softmax_temperature_fn=gomoku_visit_softmax_temperature_fn,
        known_bounds=KnownBounds(-1, 1), num_training_steps=num_training_steps, num_planes=128, num_res_blocks=8,
        hidden_dim=0, min_replay_size=min_replay_size, acc_seq_length=9999, clip_grad=clip_grad,
        use_tensorboard=use_tensorboard, is_board_game=True)


def make_classic_config(num_training_steps=100000, batch_size=256, min_replay_size=10000, use_tensorboard=True,
                        clip_grad=False) -> MuZeroConfig:
    """config.py:170-201."""
    return MuZeroConfig(
        discount=0.997, dirichlet_alpha=0.25, num_simulations=50, batch_size=batch_size, td_steps=10, lr_init=0.005,
        lr_milestones=[20000], visit_softmax_temperature_fn=classic_visit_softmax_temperature_fn,
        num_training_steps=num_training_steps, num_planes=512, num_res_blocks=0, hidden_dim=64,
        value_support_size=31, reward_support_size=31, min_replay_size=min_replay_size, checkpoint_interval=200,
        acc_seq_length=9999, clip_grad=clip_grad, use_tensorboard=use_tensorboard, is_board_game=False)


def make_atari_config(num_training_steps=int(10e6), batch_size=128, min_replay_size=10000, use_tensorboard=True,
                      clip_grad=False) -> MuZeroConfig:
    """config.py:204-233."""
    return MuZeroConfig(
        discount=0.997, dirichlet_alpha=0.25, num_simulations=30, batch_size=batch_size, td_steps=10, lr_init=0.05,
        lr_milestones=[100e3, 200e3], visit_softmax_temperature_fn=atari_visit_softmax_temperature_fn,
        num_training_steps=num_training_steps, num_planes=128, num_res_blocks=8, hidden_dim=0,
        value_support_size=61, reward_support_size=61, min_replay_size=min_replay_size, acc_seq_length=200,
        clip_grad=clip_grad, use_tensorboard=use_tensorboard, is_board_game=False)
