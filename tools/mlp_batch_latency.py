"""Search latency of the MLP configurations at small and medium batch sizes, launch chain vs the one-launch kernel.
usage: python tools/mlp_batch_latency.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import muzero_b200 as mz  # noqa: E402
from muzero_b200 import _lib  # noqa: E402
from muzero_b200.mcts import SearchPlan  # noqa: E402

for name in ('tictactoe', 'cartpole'):
    spec = bench.workload_spec(name, None)
    cfg = spec['cfg']
    net = mz.MuZeroMLPNet(**spec['net_kw'])
    net.load_state_dict(bench.state_dict_for(spec))
    net = net.cuda().eval()
    for B in (1, 32, 128, 512, 2048):
        res = {}
        for fused in (0, 1):
            plan = SearchPlan(net, cfg, B)
            eng = net.engine(B, plan.instance)
            _lib.check(_lib.lib().mz_net_set_fused_search(eng['handle'], fused))
            plan.pool.seed(1234 + np.arange(B))
            obs, mask, cur, opp = bench.synthetic_inputs(spec, B, 99)
            for _ in range(4):
                mz.uct_search_batch(obs, net, cfg, 1.0, mask, cur, opp, plan=plan)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 20
            e0.record()
            for _ in range(n):
                plan.run('device', True, False)
            e1.record()
            torch.cuda.synchronize()
            res[fused] = e0.elapsed_time(e1) / n
            _lib.check(_lib.lib().mz_net_set_fused_search(eng['handle'], 0))
            del plan
        print(f'{name} B={B}: launch chain {res[0] * 1e3:.0f} us, one-launch kernel {res[1] * 1e3:.0f} us per search '
              f'({cfg.num_simulations} simulations)', flush=True)
