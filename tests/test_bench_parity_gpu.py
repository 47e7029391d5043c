"""Parity of EXACTLY what bench.py times: the plan comes from bench.build_search -- same network
(128x8 Gomoku ResNet, seed 0), same synthetic boards, same seeds, pipeline_shape -> 2 sub-batches of 1024 trees,
tree_ctas = 4, towers capped at 144 SMs, dataflow conv launches, device-drawn Dirichlet noise, 200 simulations -- and
is run the way bench.step_device runs it (plan.run('device', True, False), eager pass + captured graph).  Then
  * >= 8 sampled trees are rebuilt bit-for-bit by the CPU oracle (visit counts, value sums, structure, action, visit
    policy, root value, RNG end state), fed the network outputs and the noise the engine used;
  * >= 64 sampled (node, action) rows of the recurrent inference (hidden state, reward, value) are compared with the
    fp32 torch restatement of the reference network."""
import os

import numpy as np
import pytest
import torch

import bench
from oracle import mcts_oracle as orc
from oracle.network_oracle import OracleNet
from parity_util import check_rows_against_oracle_net, part_of, replay_tree_in_oracle

pytestmark = pytest.mark.gpu
TOL_H, TOL_PV = 0.005, 0.02          # the stated fp16 tolerances of tests/test_conv_gpu.py


def _run_like_the_bench(workload):
    dev = torch.device('cuda', 0)
    spec = bench.workload_spec(workload, None)
    built = bench.build_search(spec, dev, rank=0)
    plan = built['plan']
    plan.run('device', True, False)          # eager pass (this call's search)
    plan.run('device', True, False)          # graph capture
    plan.run('device', True, False)          # graph replay: what the timed loop does
    torch.cuda.synchronize()
    plan.pool.check_errors()
    return spec, built


def test_gomoku_bench_plan_replays_in_the_oracle():
    spec, built = _run_like_the_bench('gomoku')
    plan = built['plan']
    # the plan bench.py builds for configs[2]
    assert hasattr(plan, 'parts') and len(plan.parts) == 2 and plan.parts[0].B == 1024
    assert plan.tree_ctas == 4 and plan.cta_limit == 144 and plan.S == 200 and plan.A == 82
    B = plan.B
    gen = np.random.RandomState(2024)
    sample = sorted(gen.choice(B, size=12, replace=False).tolist())
    # ---- trees: a fresh run from freshly seeded streams, so that ONE search separates seed and end state
    plan.pool.seed(built['seeds'])
    plan.run('device', True, False)
    torch.cuda.synchronize()
    plan.pool.check_errors()
    cfg, A = spec['cfg'], 82
    alpha = np.ones(A, dtype=np.float32) * cfg.root_dirichlet_alpha
    action, pi, rootv = plan.action.cpu(), plan.pi, plan.root_value.cpu().numpy()
    states, noise_dev = plan.pool.get_rng_states(), plan.noise.cpu().numpy()
    done = 0
    for t in sample:
        rs = np.random.RandomState(int(built['seeds'][t]))
        nz = rs.dirichlet(alpha)
        if not np.allclose(nz, noise_dev[t], rtol=1e-9, atol=1e-300):
            continue            # a rejection of the gamma sampler fell the other way by an ulp: stream use differs
        replay_tree_in_oracle(plan, t, cfg, 1.0, built['mask'][t], (built['cur'][t], built['opp'][t]), action, pi,
                              rootv, rs, noise=noise_dev[t])
        st = rs.get_state()
        assert st[2] == states[t][2] and np.array_equal(st[1], states[t][1]), f'tree {t}: RNG end state differs'
        done += 1
    assert done >= 8, f'only {done} of {len(sample)} sampled trees had numpy-identical noise'
    # ---- network rows of that same search vs the fp32 torch restatement of the reference
    sd = {k: v.detach().cpu() for k, v in built['net'].state_dict().items()}
    onet = OracleNet('board', sd, 82, 1, 1, 8)
    rows = [(int(t), int(k)) for t, k in zip(gen.choice(B, size=64), gen.randint(1, 201, size=64))]
    check_rows_against_oracle_net(built['net'], onet, plan, rows, TOL_H, TOL_PV)
    bench.release(built)


@pytest.mark.parametrize('name,trees,bound_action,bound_visits,bound_l1', [
    ('tictactoe', 128, 0.03, 0.06, 0.004), ('cartpole', 64, 0.04, 0.08, 0.004), ('gomoku', 4, 0.25, 1.0, 0.03)])
def test_fp16_networks_rarely_change_a_search(name, trees, bound_action, bound_visits, bound_l1):
    """How often a whole SEARCH changes because the engine's networks compute in fp16 on the tensor cores: the engine search
    against the CPU oracle search driven by the fp32 torch restatement of the reference network, same observation, mask
    and MT19937 stream per tree (tools/fp16_search_stats.py; profiles/r2_fp16_search_stats.json holds the figures at
    512 / 256 / 32 trees: sampled action 0 %, visit vector 0.8 % / 1.2 % / 62 %, mean L1 of the visit policy 7e-4 / 5e-4 /
    8e-3).  Bounds: fraction of trees whose sampled action / visit vector differs, mean L1 distance of the visit policies."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))
    import fp16_search_stats
    st = fp16_search_stats.search_stats(name, trees)
    assert st['sampled_action_differs'] <= bound_action, st
    assert st['visit_vector_differs'] <= bound_visits, st
    assert st['mean_l1_visit_policy'] <= bound_l1, st
