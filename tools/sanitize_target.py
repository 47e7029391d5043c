"""Small searches -- and one small training step -- for compute-sanitizer (memcheck / racecheck / synccheck): one per
kernel family.  usage: python tools/sanitize_target.py tree|conv_resident|conv_dataflow|train"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import muzero_b200 as mz  # noqa: E402

which = sys.argv[1]
dev = torch.device('cuda', 0)
torch.manual_seed(0)
gen = np.random.RandomState(0)
if which == 'train':
    # one K=3 unroll on the training kernels (csrc/train.cu): forward / dgrad / wgrad convolutions, BatchNorm kernels, tower
    # boundary kernels, stacked prediction calls (32 boards x 100 rows per call = 25 tiles), the weight-gradient streams
    from muzero_b200 import train_engine
    from muzero_b200.training import calc_loss, synthetic_transitions
    net = mz.MuZeroBoardGameNet((9, 9, 9), 82, 1, 128).to(dev).train()
    tr, w = synthetic_transitions(net, 32, 3, seed=2)
    loss, pri = calc_loss(net, dev, tr, torch.from_numpy(w).to(dev))
    eng = train_engine.engine_for(net, 32, 3)
    assert eng is not None and eng.active and eng.max_stacked_calls >= 3
    loss.backward()
    torch.cuda.synchronize()
    g = float(net.represent_net.conv_block[0].weight.grad.abs().sum())
    assert np.isfinite(float(loss)) and np.isfinite(g) and g > 0
    print('sanitize target train ok: loss %.4f' % float(loss))
    sys.exit(0)
if which == 'tree':
    # fused tree kernels (select / expand+backup / confined form) + the tcgen05 MLP kernel
    net = mz.MuZeroMLPNet((9, 3, 3), 10, 256, 1, 1, 64).to(dev).eval()
    cfg = mz.make_tictactoe_config(use_tensorboard=False)
    cfg.num_simulations = 10
    B, A = 64, 10
    obs = gen.randint(0, 2, size=(B, 9, 3, 3)).astype(np.float32)
else:
    net = mz.MuZeroBoardGameNet((9, 9, 9), 82, 2, 32).to(dev).eval()
    cfg = mz.make_gomoku_config(use_tensorboard=False)
    cfg.num_simulations = 3
    # resident launch: every CTA owns one board-aligned tile; dataflow launch: more tiles than SMs, tile flags
    B, A = (16, 82) if which == 'conv_resident' else (600, 82)
    obs = gen.randint(0, 2, size=(B, 9, 9, 9)).astype(np.float32)
mask = np.ones((B, A), dtype=bool)
plan = mz.mcts.SearchPlan(net, cfg, B)
plan.use_graph = False
if which == 'tree':
    from muzero_b200 import _lib
    _lib.check(_lib.lib().mz_pool_set_tree_ctas(plan.pool.handle, 2) if os.environ.get('MZ_SAN_CONFINED') else 0)
plan.pool.seed(np.arange(B) + 5)
a, pi, q = mz.uct_search_batch(obs, net, cfg, 1.0, mask, 1, 2, plan=plan)
torch.cuda.synchronize()
plan.pool.check_errors()
print('sanitize target', which, 'ok: actions', a[:6].tolist())
