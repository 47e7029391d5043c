"""Latency of the reference-signature call ``uct_search(state, network, device, config, ...)`` -- ONE tree, host numpy
in and out, np.random's global stream consumed like the reference does -- for the four configurations.
usage: python tools/dropin_latency.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import muzero_b200 as mz  # noqa: E402

for name in (sys.argv[1:] or ('cartpole', 'tictactoe', 'gomoku', 'atari')):
    spec = bench.workload_spec(name, 1)
    cfg = spec['cfg']
    cls = {'mlp': mz.MuZeroMLPNet, 'board': mz.MuZeroBoardGameNet, 'atari': mz.MuZeroAtariNet}[spec['kind']]
    net = cls(**spec['net_kw'])
    net.load_state_dict(bench.state_dict_for(spec))
    net = net.cuda().eval()
    obs, mask, cur, opp = bench.synthetic_inputs(spec, 1, 7)
    if hasattr(obs, 'expand'):
        obs = obs.expand().numpy()                  # the reference-signature call takes the float32 observation
    np.random.seed(0)
    dev = torch.device('cuda')
    for _ in range(5):
        mz.uct_search(obs[0], net, dev, cfg, 1.0, mask[0], int(cur[0]), int(opp[0]))
    torch.cuda.synchronize()
    n = 20
    t0 = time.perf_counter()
    for _ in range(n):
        a, pi, v = mz.uct_search(obs[0], net, dev, cfg, 1.0, mask[0], int(cur[0]), int(opp[0]))
    dt = (time.perf_counter() - t0) / n
    print(f'{name:10s} {cfg.num_simulations:4d} simulations: {dt * 1e3:8.3f} ms per uct_search call '
          f'({cfg.num_simulations / dt:9.0f} sims/s, one tree)')
