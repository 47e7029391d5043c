"""Search latency at small batch sizes (the reference's one-tree-per-actor usage) for both conv tile variants.
usage: MZ_CONV_TILE_ROWS=128|256 python tools/small_batch_latency.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import muzero_b200 as mz  # noqa: E402
from muzero_b200.mcts import SearchPlan  # noqa: E402

spec = bench.workload_spec('gomoku', None)
cfg = spec['cfg']
net = mz.MuZeroBoardGameNet(**spec['net_kw'])
net.load_state_dict(bench.state_dict_for(spec))
net = net.cuda().eval()
for B in (1, 16, 64, 256):
    plan = SearchPlan(net, cfg, B)
    plan.pool.seed(1234 + np.arange(B))
    obs, mask, cur, opp = bench.synthetic_inputs(spec, B, 99)
    for _ in range(3):
        mz.uct_search_batch(obs, net, cfg, 1.0, mask, cur, opp, plan=plan)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        mz.uct_search_batch(obs, net, cfg, 1.0, mask, cur, opp, plan=plan)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / n * 1e3
    print(f'tile_rows={os.environ.get("MZ_CONV_TILE_ROWS", "auto")} B={B}: {ms:.2f} ms per search, '
          f'{B * cfg.num_simulations / ms * 1e3:.0f} sims/s', flush=True)
