"""How often does a SEARCH change because the engine's networks compute in fp16 (tcgen05) instead of fp32?

For every tree: the engine search (numpy-exact mode: numpy draws the Dirichlet noise, the tree continues the same
MT19937 stream) against the CPU oracle search driven by the fp32 torch restatement of the reference network on the
same observation, mask and RandomState seed.  Reported per workload: fraction of trees whose sampled action / argmax
action / whole visit vector differ, and the mean L1 distance between the visit policies.

usage: python tools/fp16_search_stats.py [out.json] [trees_ttt trees_cartpole trees_gomoku]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import muzero_b200 as mz  # noqa: E402
from oracle import mcts_oracle as orc  # noqa: E402
from oracle.network_oracle import OracleNet  # noqa: E402


def search_stats(name, trees, seed0=4000):
    spec = bench.workload_spec(name, trees)
    cfg, kw = spec['cfg'], spec['net_kw']
    A, S = kw['num_actions'], cfg.num_simulations
    sd = bench.state_dict_for(spec)
    cls = {'mlp': mz.MuZeroMLPNet, 'board': mz.MuZeroBoardGameNet, 'atari': mz.MuZeroAtariNet}[spec['kind']]
    net = cls(**kw)
    net.load_state_dict(sd)
    net = net.cuda().eval()
    onet = OracleNet(spec['kind'], sd, A, kw.get('value_support_size', 1), kw.get('reward_support_size', 1),
                     kw.get('num_res_blocks', 0))
    obs, mask, cur, opp = bench.synthetic_inputs(spec, trees, 5)
    obs_f = obs.expand().numpy() if hasattr(obs, 'expand') else np.asarray(obs, dtype=np.float32)
    streams = [np.random.RandomState(seed0 + t) for t in range(trees)]
    plan = mz.mcts.SearchPlan(net, cfg, trees)
    a_e, pi_e, q_e = mz.uct_search_batch(obs, net, cfg, 1.0, mask, cur, opp, rng=streams, plan=plan)
    plan.pool.check_errors()
    a_e, pi_e, vis_e = a_e.cpu().numpy(), pi_e.cpu().numpy(), plan.visits.cpu().numpy()
    t0 = time.time()
    diff_a = diff_arg = diff_vis = 0
    l1 = []
    for t in range(trees):
        rs = np.random.RandomState(seed0 + t)
        a_o, pi_o, q_o = orc.uct_search(obs_f[t], onet, 'cpu', cfg, 1.0, mask[t], int(cur[t]), int(opp[t]), False, rng=rs)
        diff_a += int(a_o != a_e[t])
        diff_arg += int(np.argmax(pi_o) != np.argmax(pi_e[t]))
        same_vis = np.array_equal(np.round(pi_o * S).astype(np.int64), np.round(pi_e[t] * S).astype(np.int64))
        diff_vis += int(not same_vis)
        l1.append(float(np.abs(pi_o - pi_e[t]).sum()))
    out = {'workload': spec['label'], 'trees': trees, 'simulations': S, 'temperature': 1.0,
           'sampled_action_differs': diff_a / trees, 'argmax_action_differs': diff_arg / trees,
           'visit_vector_differs': diff_vis / trees, 'mean_l1_visit_policy': float(np.mean(l1)),
           'max_l1_visit_policy': float(np.max(l1)), 'oracle_seconds': time.time() - t0}
    del plan
    net.release_engine()
    return out


if __name__ == '__main__':
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'gpurun_out', 'fp16_search_stats.json')
    n = [int(x) for x in sys.argv[2:5]] if len(sys.argv) > 4 else [512, 256, 32]
    res = {}
    for name, trees in zip(('tictactoe', 'cartpole', 'gomoku'), n):
        if trees > 0:
            res[name] = search_stats(name, trees)
            print(name, json.dumps(res[name]), flush=True)
    json.dump(res, open(out_path, 'w'), indent=1)
