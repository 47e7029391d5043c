"""Helpers shared by the GPU parity tests: replay device trees in the CPU oracle (T2) and compare sampled network rows
of a finished search with the fp32 torch restatement of the reference network."""
import numpy as np
import torch

from conftest import bits
from oracle import mcts_oracle as orc
from oracle.stubnet import ReplayStub


def part_of(plan, t):
    """(SearchPlan, local tree index) that owns tree t of a plan (pipelined plans split the batch in parts)."""
    if hasattr(plan, 'parts'):
        per = plan.parts[0].B
        return plan.parts[t // per], t % per
    return plan, t


def replay_tree_in_oracle(plan, t, cfg, temperature, mask_t, players, action, pi, rootv, rng, noise=None,
                          deterministic=False):
    """Tree t, fed the per-node (reward, value) the engine produced and the noise it used, must be rebuilt bit-for-bit
    by the oracle: visit counts, value sums, chosen action, visit policy, root value; returns the oracle's trace."""
    d = plan.pool.dump_tree(t)
    pi0 = plan.pi0[t].cpu().numpy()
    stub = ReplayStub(pi0, d['R'][1:].astype(np.float32), d['value'][1:], d['parent'], d['move'])
    a_o, pi_o, q_o, tr = orc.uct_search(np.zeros(1, np.float32), stub, 'cpu', cfg, temperature, mask_t, int(players[0]),
                                        int(players[1]), deterministic, rng=rng, noise=noise, return_trace=True)
    assert np.array_equal(tr.N, d['N']), f'tree {t}: visit counts differ'
    assert np.array_equal(bits(tr.W), bits(d['W'])), f'tree {t}: value sums differ'
    assert np.array_equal(tr.parent, d['parent']) and np.array_equal(tr.move, d['move']), f'tree {t}: structure differs'
    assert a_o == int(action[t]), f'tree {t}: action {int(action[t])} vs oracle {a_o}'
    assert np.array_equal(bits(pi_o), bits(pi[t].cpu().numpy())), f'tree {t}: visit policy differs'
    assert bits(q_o)[0] == bits(float(rootv[t]))[0], f'tree {t}: root value differs'
    return d, tr


def check_rows_against_oracle_net(net, onet, plan, rows, tol_h, tol_pv):
    """rows: list of (tree, node >= 1).  The child's hidden state / reward / value the engine stored for that node
    against OracleNet.recurrent_batch on the PARENT's stored hidden state and the node's action (fp32 torch)."""
    hs, acts, got_h, got_r, got_v = [], [], [], [], []
    cache = {}
    for t, k in rows:
        part, lt = part_of(plan, t)
        if t not in cache:
            cache[t] = plan.pool.dump_tree(t)
        d = cache[t]
        S1 = part.S + 1
        slots = part.pool.hidden.view(part.B, S1, -1)[lt]
        par = int(d['parent'][k])
        pair = net.hidden_to_reference(slots[[par, k]]).cpu()
        hs.append(pair[0]); got_h.append(pair[1].numpy())
        acts.append(int(d['move'][k])); got_r.append(float(d['R'][k])); got_v.append(float(d['value'][k]))
    h2, r, _, v = onet.recurrent_batch(torch.stack(hs), np.array(acts))
    got_h, h2 = np.stack(got_h), h2.numpy()
    eh = float(np.abs(got_h - h2).max())
    sr, sv = max(1.0, float(np.abs(r.numpy()).max())), max(1.0, float(np.abs(v.numpy()).max()))
    er = float(np.abs(np.array(got_r) - r.numpy()).max())
    ev = float(np.abs(np.array(got_v) - v.numpy()).max())
    print(f'{len(rows)} sampled (node, action) rows vs fp32 torch: hidden {eh:.4g} (tol {tol_h}), reward {er:.4g} '
          f'(tol {tol_pv * sr:.4g}), value {ev:.4g} (tol {tol_pv * sv:.4g})')
    assert np.isfinite(got_h).all() and eh <= tol_h
    assert er <= tol_pv * sr and ev <= tol_pv * sv
