"""GPU parity tests of the tree kernels, through the C ABI (SearchPool is a thin
ctypes wrapper) and through the public ``uct_search``.  Bar: bit-exact."""
import zlib

import numpy as np
import pytest
import torch

from conftest import bits, load_golden_cases
from oracle import mcts_oracle as orc
from oracle.stubnet import HashStub, ReplayStub

pytestmark = pytest.mark.gpu
CASES = load_golden_cases()


def _dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    return torch.device('cuda', 0)


def _rng_sig(state):
    return int(state[2]), zlib.crc32(np.asarray(state[1], np.uint32).tobytes())


@pytest.mark.parametrize('case', CASES, ids=[c.name for c in CASES])
def test_golden_replay_through_c_abi(case):
    """Kernels fed the network outputs and noise the reference saw must rebuild the
    reference's tree, policy, action, root value and RNG stream bit-for-bit."""
    import muzero_b200 as mz
    dev = _dev()
    cfg = case.config()
    pool = mz.SearchPool(1, case.A, cfg, 0, dev)
    np.random.seed(case.seed)
    noise = None
    use_noise = (not case.deterministic) and case.alpha > 0 and case.eps > 0
    if use_noise:
        noise = torch.from_numpy(np.random.dirichlet(np.ones_like(case.root_pi) * case.alpha)[None]).to(dev)
    pool.set_rng_states([np.random.get_state()])
    mask = None if case.mask is None else torch.from_numpy(case.mask.astype(np.uint8)[None].copy()).to(dev)
    players = torch.tensor([list(case.players)], dtype=torch.int32, device=dev)
    pool.reset(torch.from_numpy(case.root_pi[None].copy()).to(dev), noise, case.eps if use_noise else 0.0, mask,
               players)
    rew = torch.from_numpy(case.rewards.copy()).to(dev)
    val = torch.from_numpy(case.values.copy()).to(dev)
    got_parent, got_action = [], []
    for i in range(case.sims):
        pool.select()
        got_parent.append(pool.view('LEAF_PARENT').clone())
        got_action.append(pool.view('LEAF_ACTION').clone())
        pool.expand_backup(rew[i:i + 1], val[i:i + 1])
    temps = torch.tensor([case.temperature], dtype=torch.float64, device=dev)
    action, pi, rootv, visits = pool.root_policy(mask, temps, case.deterministic)
    pool.check_errors()
    assert np.array_equal(torch.cat(got_parent).cpu().numpy(), case.parent[1:])
    assert np.array_equal(torch.cat(got_action).cpu().numpy(), case.move[1:])
    tree = pool.dump_tree(0)
    assert tree['num_nodes'] == case.sims + 1
    assert np.array_equal(tree['N'], case.N)
    assert np.array_equal(bits(tree['W']), bits(case.W))
    assert np.array_equal(bits(tree['R']), bits(case.R))
    assert np.array_equal(tree['parent'], case.parent) and np.array_equal(tree['move'], case.move)
    assert np.array_equal(bits(tree['prior']), bits(case.prior.astype(np.float64)))
    assert int(action[0].cpu()) == case.action
    assert np.array_equal(bits(pi[0].cpu().numpy()), bits(case.pi))
    assert bits(float(rootv[0].cpu()))[0] == bits(case.root_value)[0]
    assert _rng_sig(pool.get_rng_states()[0]) == case.rng_end


@pytest.mark.parametrize('case', CASES[::3], ids=[c.name for c in CASES[::3]])
def test_public_uct_search_with_foreign_network(case):
    """The drop-in entry point with a duck-typed network object (the reference accepts any
    object with the two inference methods) and numpy's global stream."""
    import muzero_b200 as mz
    np.random.seed(case.seed)
    net = ReplayStub(case.root_pi, case.rewards, case.values, case.parent, case.move)
    a, pi, q = mz.uct_search(np.zeros((2, 2), np.float32), net, 'cpu', case.config(), case.temperature, case.mask,
                             case.players[0], case.players[1], case.deterministic)
    assert isinstance(a, int) and isinstance(q, float) and pi.dtype == np.float64 and pi.shape == (case.A,)
    assert a == case.action
    assert np.array_equal(bits(pi), bits(case.pi))
    assert bits(q)[0] == bits(case.root_value)[0]
    assert _rng_sig(np.random.get_state()) == case.rng_end


def _lockstep(B, A, S, board, bounds, discount, alpha, noise_on, seed, quantise=0, value_scale=1.0,
              reward_scale=0.0, temperature=1.0, deterministic=False, uniform_prior=False):
    """Run B device trees and B oracle trees side by side, comparing after every step."""
    import muzero_b200 as mz
    dev = _dev()
    gen = np.random.RandomState(seed)
    cfg = mz.MuZeroConfig(discount=discount, dirichlet_alpha=alpha, num_simulations=S, batch_size=1, td_steps=0,
                          lr_init=0.0, lr_milestones=[], visit_softmax_temperature_fn=None,
                          known_bounds=mz.KnownBounds(-1, 1) if bounds else None, is_board_game=board)
    logits = np.zeros((B, A)) if uniform_prior else gen.standard_normal((B, A)) * 2
    e = np.exp(logits - logits.max(1, keepdims=True))
    pi0 = (e / e.sum(1, keepdims=True)).astype(np.float32)
    mask = gen.rand(B, A) < 0.7
    mask[np.arange(B), gen.randint(A, size=B)] = True
    players = np.stack([gen.randint(1, 3, size=B), np.zeros(B, int)], 1)
    players[:, 1] = np.where(board, 3 - players[:, 0], players[:, 0])
    streams = [np.random.RandomState(int(s)) for s in gen.randint(1 << 30, size=B)]
    noise = None
    if noise_on:
        noise = np.stack([r.dirichlet(np.ones(A, np.float32) * alpha) for r in streams])
    pool = mz.SearchPool(B, A, cfg, 0, dev)
    pool.set_rng_states([r.get_state() for r in streams])
    pool.reset(torch.from_numpy(pi0).to(dev), None if noise is None else torch.from_numpy(noise).to(dev),
               cfg.root_exploration_eps if noise_on else 0.0, torch.from_numpy(mask.astype(np.uint8)).to(dev),
               torch.from_numpy(players.astype(np.int32)).to(dev))
    trees, stubs = [], []
    for t in range(B):
        p = pi0[t]
        if noise_on:
            p = orc.mix_dirichlet(p, noise[t], cfg.root_exploration_eps)
        p = orc.mask_and_renormalise(mask[t], p)
        tr = orc.OracleTree(A, S, p, discount, board, cfg.known_bounds, cfg.pb_c_base, cfg.pb_c_init,
                            int(players[t, 0]), int(players[t, 1]), orc.MT19937.from_numpy(streams[t]))
        stub = HashStub(pi0[t], value_scale, reward_scale, quantise, seed=seed * 1000 + t)
        o = stub.initial_inference(None)
        tr.set_root(o.hidden_state, 0.0)
        trees.append(tr); stubs.append(stub)
    assert np.array_equal(bits(pool.view('PRIOR').view(B, A).cpu().numpy()),
                          bits(np.stack([tr.prior.astype(np.float64) for tr in trees])).reshape(B, A))
    for i in range(S):
        pool.select()
        gp = pool.view('LEAF_PARENT').cpu().numpy(); ga = pool.view('LEAF_ACTION').cpu().numpy()
        gd = pool.view('LEAF_DEPTH').cpu().numpy()
        r = np.zeros(B, np.float32); v = np.zeros(B, np.float32)
        for t in range(B):
            n, a, pl, d = trees[t].select()
            assert (gp[t], ga[t], gd[t]) == (n, a, d), f'sim {i} tree {t}: device {(gp[t], ga[t], gd[t])} oracle {(n, a, d)}'
            o = stubs[t].recurrent_inference(trees[t].hidden[n], np.array([a]))
            trees[t].expand_backup(n, a, pl, d, o.hidden_state, o.reward, o.value)
            r[t], v[t] = o.reward, o.value
        pool.expand_backup(torch.from_numpy(r).to(dev), torch.from_numpy(v).to(dev))
    mm = pool.view('MINMAX').view(B, 2).cpu().numpy()
    temps = torch.full((B,), temperature, dtype=torch.float64, device=dev)
    action, pi, rootv, visits = pool.root_policy(torch.from_numpy(mask.astype(np.uint8)).to(dev), temps, deterministic)
    pool.check_errors()
    action, pi, rootv, visits = action.cpu().numpy(), pi.cpu().numpy(), rootv.cpu().numpy(), visits.cpu().numpy()
    states = pool.get_rng_states()
    for t in range(B):
        tr = trees[t]
        d = pool.dump_tree(t)
        assert np.array_equal(d['N'], tr.Nv) and np.array_equal(bits(d['W']), bits(tr.W))
        assert np.array_equal(bits(d['R']), bits(tr.R))
        assert np.array_equal(d['parent'], tr.PAR) and np.array_equal(d['move'], tr.MOVE)
        assert np.array_equal(d['children'], tr.CH)
        assert bits(mm[t, 0])[0] == bits(tr.lo)[0] and bits(mm[t, 1])[0] == bits(tr.hi)[0]
        vis = np.where(mask[t], tr.root_visits(), 0)
        assert np.array_equal(visits[t], vis)
        want_pi = orc.play_policy(vis, temperature)
        assert np.array_equal(bits(pi[t]), bits(want_pi))
        want_a = int(np.argmax(vis)) if deterministic else tr.rng.choice_p(want_pi)
        assert action[t] == want_a
        assert bits(rootv[t])[0] == bits(tr.W[0] / int(tr.Nv[0]))[0]
        assert states[t][2] == tr.rng.pos and np.array_equal(states[t][1], tr.rng.key)


@pytest.mark.parametrize('kw', [
    dict(B=33, A=10, S=25, board=True, bounds=True, discount=1.0, alpha=0.25, noise_on=True, seed=1),
    dict(B=16, A=2, S=50, board=False, bounds=False, discount=0.997, alpha=0.25, noise_on=False, seed=2,
         reward_scale=1.0, deterministic=True, temperature=0.25),
    dict(B=8, A=82, S=60, board=True, bounds=True, discount=1.0, alpha=0.03, noise_on=True, seed=3, value_scale=3.0),
    dict(B=8, A=18, S=40, board=False, bounds=False, discount=0.997, alpha=0.25, noise_on=True, seed=4,
         reward_scale=2.0, temperature=0.5),
    # exact ties everywhere: uniform prior, no noise, coarse values -> the MT19937 tie-break path, incl. a twist
    dict(B=8, A=10, S=120, board=True, bounds=True, discount=1.0, alpha=0.25, noise_on=False, seed=5, quantise=2,
         uniform_prior=True, temperature=0.1),
    dict(B=4, A=226, S=40, board=True, bounds=True, discount=1.0, alpha=0.03, noise_on=True, seed=6),
    dict(B=5, A=1, S=10, board=False, bounds=False, discount=0.9, alpha=0.25, noise_on=False, seed=7),
    # >= 512 trees with <= 4 actions: the thread-per-tree kernels (no bounds: the child_Q cache is refreshed wholesale;
    # uniform prior + coarse values: ties, i.e. the sequential MT19937 path)
    dict(B=520, A=2, S=14, board=False, bounds=False, discount=0.997, alpha=0.25, noise_on=True, seed=8,
         reward_scale=1.0),
    dict(B=512, A=4, S=12, board=True, bounds=True, discount=1.0, alpha=0.25, noise_on=False, seed=9, quantise=2,
         uniform_prior=True),
], ids=['ttt', 'cartpole', 'gomoku', 'atari', 'ties', 'gomoku15', 'single-action', 'thread-per-tree-A2',
        'thread-per-tree-A4-ties'])
def test_batched_trees_lockstep_with_oracle(kw):
    _lockstep(**kw)


def _synthetic_outputs(B, S, dev, seed):
    """Per-(sim, tree) reward/value streams, float32, generated on the device."""
    g = torch.Generator(device=dev).manual_seed(seed)
    value = (torch.rand((S, B), generator=g, device=dev) * 2 - 1).contiguous()
    reward = torch.zeros((S, B), device=dev)
    return reward, value


@pytest.mark.parametrize('B,A,S,alpha', [(4096, 10, 25, 0.25), (2048, 82, 200, 0.03)], ids=['C2-size', 'C3-size'])
def test_full_size_properties_and_sampled_replay(B, A, S, alpha):
    """BASELINE.json sizes: structural invariants on every tree, plus a sample of trees
    replayed through the oracle bit-for-bit."""
    import muzero_b200 as mz
    dev = _dev()
    cfg = mz.MuZeroConfig(discount=1.0, dirichlet_alpha=alpha, num_simulations=S, batch_size=1, td_steps=0,
                          lr_init=0.0, lr_milestones=[], visit_softmax_temperature_fn=None,
                          known_bounds=mz.KnownBounds(-1, 1), is_board_game=True)
    gen = np.random.RandomState(B + A)
    logits = gen.standard_normal((B, A)).astype(np.float32)
    pi0 = torch.softmax(torch.from_numpy(logits), dim=1).numpy()
    mask = gen.rand(B, A) < 0.8
    mask[:, -1] = True
    pool = mz.SearchPool(B, A, cfg, 0, dev)
    seeds = np.arange(B) + 1234
    pool.seed(seeds)
    sample = sorted(gen.choice(B, size=6, replace=False).tolist())
    # numpy-drawn noise for the sampled trees (bit-exact path), zeros-safe noise elsewhere is fine too:
    noise = np.stack([np.random.RandomState(int(s)).dirichlet(np.ones(A, np.float32) * alpha) for s in seeds[:64]])
    noise = np.concatenate([noise] * (B // 64 + 1))[:B]
    streams = {}
    for t in sample:
        rs = np.random.RandomState(int(seeds[t]))
        noise[t] = rs.dirichlet(np.ones(A, np.float32) * alpha)
        streams[t] = rs
    st = pool.get_rng_states()
    for t in sample:                                      # sampled trees continue their numpy stream after the noise
        st[t] = streams[t].get_state()
    pool.set_rng_states(st)
    mask_d = torch.from_numpy(mask.astype(np.uint8)).to(dev)
    players = torch.tensor([[1, 2]] * B, dtype=torch.int32, device=dev)
    pool.reset(torch.from_numpy(pi0).to(dev), torch.from_numpy(noise).to(dev), 0.25, mask_d, players)
    reward, value = _synthetic_outputs(B, S, dev, 7)
    parents = torch.empty((S, B), dtype=torch.int32, device=dev)
    actions = torch.empty((S, B), dtype=torch.int32, device=dev)
    for i in range(S):
        pool.select()
        parents[i].copy_(pool.view('LEAF_PARENT')); actions[i].copy_(pool.view('LEAF_ACTION'))
        pool.expand_backup(reward[i], value[i])
    temps = torch.ones(B, dtype=torch.float64, device=dev)
    action, pi, rootv, visits = pool.root_policy(mask_d, temps, False)
    pool.check_errors()
    # --- invariants over all trees
    assert (pool.view('COUNT') == S + 1).all() and (pool.view('ROOT_N') == S).all()
    hot = pool.view('EDGES').view(B, (S + 1) * A, 8).cpu().numpy().view(
        np.dtype([('N', '<u2'), ('child', '<u2'), ('Q', '<f4')])).reshape(B, S + 1, A)
    rec = {'N': hot['N'], 'child': hot['child'],
           'W': np.where(hot['N'] > 0, pool.view('EDGE_W').view(B, S + 1, A).cpu().numpy(), 0.0)}
    root_child_visits = rec['N'][:, 0, :].astype(np.int64)
    assert (root_child_visits.sum(1) == S).all()                      # every simulation passes one root child
    has_child = rec['child'] != 0xFFFF
    assert (has_child.sum((1, 2)) == S).all()                         # S expansions -> S child links
    assert ((rec['N'] > 0) == has_child).all()                        # visited <=> expanded
    child_sum = rec['N'].astype(np.int64).sum(2)                      # per node: sum of children visits
    par = pool.view('NODE_PARENT').view(B, S + 1).cpu().numpy(); mv = pool.view('NODE_MOVE').view(B, S + 1).cpu().numpy()
    tt = np.arange(B)[:, None].repeat(S, 1)
    node_N = rec['N'][tt, par[:, 1:], mv[:, 1:]].astype(np.int64)     # N of nodes 1..S
    assert (node_N == 1 + child_sum[:, 1:]).all()                     # N(node) = 1 + sum N(children)
    assert (np.abs(rec['W'][tt, par[:, 1:], mv[:, 1:]]) <= node_N + 1e-9).all()   # |value| <= 1 each
    v = visits.cpu().numpy()
    assert (v == np.where(mask, root_child_visits, 0)).all()
    p = pi.cpu().numpy()
    assert np.allclose(p.sum(1), 1.0) and (p[~mask] == 0).all()
    a = action.cpu().numpy()
    assert ((a >= 0) & (a < A)).all() and mask[np.arange(B), a].all() and (v[np.arange(B), a] > 0).all()
    # --- sampled trees: oracle replay, bit-exact
    parents, actions = parents.cpu().numpy(), actions.cpu().numpy()
    reward, value = reward.cpu().numpy(), value.cpu().numpy()
    states = pool.get_rng_states()
    for t in sample:
        rs = np.random.RandomState(int(seeds[t]))
        par_t = np.concatenate([[-1], parents[:, t]]); mv_t = np.concatenate([[-1], actions[:, t]])
        net = ReplayStub(pi0[t], reward[:, t], value[:, t], par_t, mv_t)
        a_o, pi_o, q_o, tr = orc.uct_search(np.zeros(1, np.float32), net, 'cpu', cfg, 1.0, mask[t], 1, 2, False,
                                            rng=rs, return_trace=True)
        d = pool.dump_tree(t)
        assert np.array_equal(d['N'], tr.N) and np.array_equal(bits(d['W']), bits(tr.W))
        assert a_o == a[t] and np.array_equal(bits(pi_o), bits(p[t]))
        assert bits(q_o)[0] == bits(rootv[t].item())[0]
        assert _rng_sig(states[t]) == _rng_sig(rs.get_state())


def test_device_seed_matches_numpy_seed():
    import muzero_b200 as mz
    dev = _dev()
    cfg = mz.make_tictactoe_config(use_tensorboard=False)
    pool = mz.SearchPool(5, 10, cfg, 0, dev)
    seeds = [0, 1, 1234, 2**31 + 5, 2**32 - 1]
    pool.seed(seeds)
    for s, st in zip(seeds, pool.get_rng_states()):
        want = np.random.RandomState(s).get_state()
        assert np.array_equal(st[1], want[1]) and st[2] == want[2]


def test_device_dirichlet_follows_numpy_to_ulps():
    """Same algorithm and draws as numpy's legacy sampler; log/pow are CUDA's, so the
    contract is agreement to a few ulp and identical stream consumption — parity tests
    inject numpy's noise instead."""
    import muzero_b200 as mz
    dev = _dev()
    cfg = mz.make_tictactoe_config(use_tensorboard=False)
    B, A = 64, 10
    pool = mz.SearchPool(B, A, cfg, 0, dev)
    pool.seed(np.arange(B) + 99)
    out = pool.dirichlet(0.25).cpu().numpy()
    states = pool.get_rng_states()
    same_stream = 0
    for t in range(B):
        rs = np.random.RandomState(99 + t)
        want = rs.dirichlet(np.ones(A) * 0.25)
        if rs.get_state()[2] == states[t][2]:
            same_stream += 1
            assert np.allclose(out[t], want, rtol=1e-12, atol=1e-300)
        assert abs(out[t].sum() - 1) < 1e-12 and (out[t] >= 0).all()
    assert same_stream >= B - 2        # a rejection decided differently by an ulp is possible but rare


def test_error_paths_match_reference_exceptions():
    import muzero_b200 as mz
    dev = _dev()
    cfg = mz.make_tictactoe_config(use_tensorboard=False)
    pool = mz.SearchPool(2, 10, cfg, 0, dev)
    with pytest.raises(Exception, match='without a preceding mz_select'):
        pool.expand_backup(torch.zeros(2, device=dev), torch.zeros(2, device=dev))
    with pytest.raises(ValueError, match='Expect `temperature`'):
        mz.uct_search(np.zeros((2, 2), np.float32), ReplayStub(np.ones(10, np.float32) / 10, [], []), 'cpu', cfg, 3.0,
                      np.ones(10, bool), 1, 2)
    bad = mz.make_tictactoe_config(use_tensorboard=False)
    bad.discount = 0.9
    with pytest.raises(AssertionError):
        mz.SearchPool(2, 10, bad, 0, dev)
    # all visits masked away -> the reference's np.random.choice raises on the NaN policy
    pool.seed([1, 2])
    pi0 = torch.full((2, 10), 0.1, device=dev)
    pool.reset(pi0, None, 0.0, None, None)
    for _ in range(cfg.num_simulations):
        pool.select()
        pool.expand_backup(torch.zeros(2, device=dev), torch.zeros(2, device=dev))
    none = torch.zeros((2, 10), dtype=torch.uint8, device=dev)
    pool.root_policy(none, torch.ones(2, dtype=torch.float64, device=dev), False)
    with pytest.raises(ValueError, match='NaN'):
        pool.check_errors()


@pytest.mark.parametrize('A,S,board,bounds', [(2, 50, False, False), (10, 25, True, True), (40, 60, False, True),
                                              (82, 120, True, True), (120, 40, False, False)],
                         ids=['A2', 'A10', 'A40', 'A82', 'A120'])
def test_fused_and_confined_tree_kernels_equal_the_separate_launches(A, S, board, bounds):
    """mz_expand_backup_select (one warp per tree over all SMs) and its confined form (mz_pool_set_tree_ctas: a few
    persistent CTAs pulling trees off a counter) against mz_expand_backup + mz_select, which the lock-step tests pin
    to the oracle: every byte of the search state, the child_Q cache and the RNG streams included."""
    import muzero_b200 as mz
    from muzero_b200 import _lib
    dev = _dev()
    lib = _lib.lib()
    B = 300
    gen = np.random.RandomState(A * 1000 + S)
    cfg = mz.MuZeroConfig(discount=1.0 if board else 0.997, dirichlet_alpha=0.25, num_simulations=S, batch_size=1,
                          td_steps=0, lr_init=0.0, lr_milestones=[], visit_softmax_temperature_fn=None,
                          known_bounds=mz.KnownBounds(-1, 1) if bounds else None, is_board_game=board)
    pi0 = gen.dirichlet(np.ones(A), size=B).astype(np.float32)
    pi0[::3] = 1.0 / A                                    # tie-heavy trees: the RNG streams matter
    noise = gen.dirichlet(np.ones(A) * 0.25, size=B)
    players = np.stack([np.ones(B), 2 * np.ones(B) if board else np.ones(B)], 1).astype(np.int32)
    rewards = (gen.standard_normal((S, B)) * (0.0 if board else 0.5)).astype(np.float32)
    values = np.tanh(gen.standard_normal((S, B))).astype(np.float32) * (1.0 if bounds else 3.0)
    values[:, ::5] = np.round(values[:, ::5])              # exact zeros / +-1: W == 0 takes the IEEE division path

    def run(mode):
        pool = mz.SearchPool(B, A, cfg, 0, dev)
        pool.seed(77 + np.arange(B))
        pool.reset(torch.from_numpy(pi0).to(dev), torch.from_numpy(noise).to(dev), 0.25, None,
                   torch.from_numpy(players).to(dev))
        if mode == 'confined':
            _lib.check(lib.mz_pool_set_tree_ctas(pool.handle, 3))
        st = _lib.current_stream()
        pool.select()
        for i in range(S):
            r, v = torch.from_numpy(rewards[i]).to(dev), torch.from_numpy(values[i]).to(dev)
            if mode == 'separate':
                pool.expand_backup(r, v)
                if i + 1 < S:
                    pool.select()
            elif i + 1 < S:
                _lib.check(lib.mz_expand_backup_select(pool.handle, r.data_ptr(), v.data_ptr(), st))
            else:
                pool.expand_backup(r, v)
            torch.cuda.synchronize()
        pool.check_errors()
        names = ('EDGES', 'EDGE_W', 'EDGE_REWARD', 'MINMAX', 'ROOT_W', 'ROOT_N', 'COUNT', 'NODE_PARENT', 'NODE_MOVE',
                 'RNG_KEY', 'RNG_POS', 'PRIOR')
        return {k: pool.view(k).cpu().numpy().view(np.uint8).copy() for k in names}

    want = run('separate')
    for mode in ('fused', 'confined'):
        got = run(mode)
        for k in want:
            assert np.array_equal(want[k], got[k]), f'{mode}: {k} differs'


def test_fresh_pools_are_seeded_and_the_default_batch_call_is_not_degenerate():
    """mz_pool_create zeroes the MT19937 states and an all-zero state emits 0 forever (NaN Dirichlet noise, constant
    tie-breaks).  SearchPool seeds every new pool from OS entropy, so uct_search_batch(rng=None) on a plan the caller
    never touched -- the call INTEGRATION.md shows -- must search with proper noise and distinct streams."""
    import muzero_b200 as mz
    dev = _dev()
    cfg = mz.make_tictactoe_config(use_tensorboard=False)
    pool = mz.SearchPool(8, 10, cfg, 0, dev)
    keys = pool.view('RNG_KEY').view(8, 624).cpu().numpy()
    assert (keys != 0).any(axis=1).all() and len({k.tobytes() for k in keys}) == 8
    net = mz.MuZeroMLPNet((9, 3, 3), 10, 256, 1, 1, 64).to(dev).eval()
    mz.mcts._PLANS.clear()
    obs = np.random.RandomState(0).randint(0, 2, size=(32, 9, 3, 3)).astype(np.float32)
    a, pi, q = mz.uct_search_batch(obs, net, cfg, 1.0, np.ones((32, 10), bool), 1, 2)
    plan = mz.mcts._plan_for(net, cfg, 32)
    plan.pool.check_errors()
    prior = plan.pool.view('PRIOR').view(32, 10).cpu().numpy()
    assert np.isfinite(prior).all() and np.allclose(prior.sum(1), 1.0)
    assert np.isfinite(pi.cpu().numpy()).all() and np.allclose(pi.cpu().numpy().sum(1), 1.0)
    assert len({p.tobytes() for p in plan.noise.cpu().numpy()}) == 32      # 32 different Dirichlet samples
    mz.mcts._PLANS.clear()
