"""muzero_b200 — B200-native MuZero search-and-inference engine.

Drop-in for the self-play hot path of michaelnny/muzero: ``uct_search`` and the
``MuZeroNet`` inference methods keep their signatures; the work runs in
hand-written sm_100a CUDA behind the C ABI of ``include/muzero_b200.h``.
"""
from .config import (KnownBounds, MuZeroConfig, make_atari_config, make_classic_config, make_gomoku_config,
                     make_tictactoe_config)
from .network import MuZeroAtariNet, MuZeroBoardGameNet, MuZeroMLPNet, MuZeroNet, NetworkOutputs, StackedFrames
from .mcts import SearchPool, uct_search, uct_search_batch
from .selfplay import BatchedBoardEnv, BoardSelfPlay, mc_return_targets, n_step_targets, unroll_sequences
from .replay import DeviceReplay

__all__ = ['KnownBounds', 'MuZeroConfig', 'make_atari_config', 'make_classic_config', 'make_gomoku_config',
           'make_tictactoe_config', 'MuZeroAtariNet', 'MuZeroBoardGameNet', 'MuZeroMLPNet', 'MuZeroNet',
           'NetworkOutputs', 'StackedFrames', 'SearchPool', 'uct_search', 'uct_search_batch', 'BatchedBoardEnv', 'BoardSelfPlay',
           'mc_return_targets', 'n_step_targets', 'unroll_sequences', 'DeviceReplay']
