"""Short run of the hot path for ncu: a few simulations of one batched search.
usage: python tools/profile_target.py [workload] [simulations] [trees]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import muzero_b200 as mz  # noqa: E402
from muzero_b200.mcts import SearchPlan  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'gomoku'
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 4
spec = bench.workload_spec(name, int(sys.argv[3]) if len(sys.argv) > 3 else None)
cfg = spec['cfg']
cfg.num_simulations = sims
cls = {'mlp': mz.MuZeroMLPNet, 'board': mz.MuZeroBoardGameNet, 'atari': mz.MuZeroAtariNet}[spec['kind']]
net = cls(**spec['net_kw'])
net.load_state_dict(bench.state_dict_for(spec))
net = net.cuda().eval()
B = spec['trees']
plan = SearchPlan(net, cfg, B)
plan.use_graph = False
plan.pool.seed(1234 + np.arange(B))
obs, mask, cur, opp = bench.synthetic_inputs(spec, B, 99)
for _ in range(2):
    mz.uct_search_batch(obs, net, cfg, 1.0, mask, cur, opp, plan=plan)
torch.cuda.synchronize()
print('done', name, B, sims)
