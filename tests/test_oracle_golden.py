"""CPU tests: the oracle against the committed reference recordings, and the
restated numpy internals (MT19937 stream, pairwise sum, Dirichlet, choice)
against numpy itself."""
import zlib

import numpy as np
import pytest

from conftest import bits, load_golden_cases
from oracle import mcts_oracle as orc
from oracle.stubnet import ReplayStub

CASES = load_golden_cases()


@pytest.mark.parametrize('case', CASES, ids=[c.name for c in CASES])
def test_oracle_reproduces_reference_recording(case):
    """Feed the oracle the network outputs the reference saw; everything the reference
    produced (action, pi, root value, every node's N/W/reward/parent/move, the prior,
    the final MT19937 state) must come out bit-for-bit."""
    np.random.seed(case.seed)
    net = ReplayStub(case.root_pi, case.rewards, case.values, case.parent, case.move)
    a, pi, q, tr = orc.uct_search(np.zeros((2, 2), np.float32), net, 'cpu', case.config(), case.temperature,
                                  case.mask, case.players[0], case.players[1], case.deterministic,
                                  return_trace=True)
    assert a == case.action
    assert np.array_equal(bits(pi), bits(case.pi))
    assert bits(q)[0] == bits(case.root_value)[0]
    assert tr.num_nodes == len(case.N) == case.sims + 1
    assert np.array_equal(tr.N, case.N) and np.array_equal(tr.parent, case.parent) and np.array_equal(tr.move, case.move)
    assert np.array_equal(bits(tr.W), bits(case.W)) and np.array_equal(bits(tr.R), bits(case.R))
    assert tr.prior.dtype == case.prior.dtype and np.array_equal(tr.prior.view(np.uint8), case.prior.view(np.uint8))
    st = np.random.get_state()
    assert (int(st[2]), zlib.crc32(np.asarray(st[1], np.uint32).tobytes())) == case.rng_end


def test_golden_covers_the_interesting_regimes():
    assert any(c.board for c in CASES) and any(not c.board for c in CASES)
    assert any(c.deterministic for c in CASES) and any(not c.deterministic for c in CASES)
    assert {c.A for c in CASES} >= {2, 10, 18, 82} and {c.sims for c in CASES} >= {25, 50, 200}
    assert any(c.mask is not None and not c.mask.all() for c in CASES) and any(c.mask is None for c in CASES)
    assert any(c.prior.dtype == np.float32 for c in CASES) and any(c.prior.dtype == np.float64 for c in CASES)
    assert {c.temperature for c in CASES} >= {0.0, 0.1, 0.25, 0.5, 1.0}
    assert any(np.abs(c.W).max() > c.sims for c in CASES)        # values outside known_bounds


def test_mt19937_matches_numpy_stream():
    for seed in (0, 1, 12345, 2**32 - 1):
        rs = np.random.RandomState(seed)
        mt = orc.MT19937.from_seed(seed)
        assert np.array_equal(mt.key, rs.get_state()[1]) and mt.pos == rs.get_state()[2]
        want = rs.randint(0, 2**32, size=2000, dtype=np.uint64)
        got = np.array([mt.next_u32() for _ in range(2000)], dtype=np.uint64)
        assert np.array_equal(want, got)
        assert [mt.next_double() for _ in range(50)] == list(rs.random_sample(50))


@pytest.mark.parametrize('k', [1, 2, 3, 5, 7, 10, 18, 33, 82, 226])
def test_tie_break_matches_numpy_choice(k):
    rs = np.random.RandomState(99 + k)
    mt = orc.MT19937.from_numpy(rs)
    ties = np.arange(k) * 3
    for _ in range(200):
        assert int(rs.choice(ties)) == int(ties[mt.bounded(k)])
    assert rs.get_state()[2] == mt.pos and np.array_equal(rs.get_state()[1], mt.key)


@pytest.mark.parametrize('alpha', [0.03, 0.25, 1.0, float(np.float32(0.03))])
@pytest.mark.parametrize('A', [2, 10, 82])
def test_dirichlet_matches_numpy(alpha, A):
    rs = np.random.RandomState(4242)
    mt = orc.MT19937.from_numpy(rs)
    for _ in range(20):
        want = rs.dirichlet(np.ones(A) * alpha)
        got = mt.dirichlet(np.ones(A) * alpha)
        assert np.array_equal(bits(want), bits(got))
    assert rs.get_state()[2] == mt.pos


def test_choice_p_matches_numpy():
    rs = np.random.RandomState(5)
    mt = orc.MT19937.from_numpy(rs)
    gen = np.random.RandomState(6)
    for _ in range(300):
        A = int(gen.choice([2, 10, 82]))
        v = gen.randint(0, 30, size=A).astype(np.float64)
        v[gen.randint(A)] += 1
        p = v / v.sum()
        assert int(rs.choice(np.arange(A), p=p)) == mt.choice_p(p)


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('n', [1, 2, 7, 8, 9, 10, 18, 82, 127, 128, 129, 226, 362, 1000])
def test_pairwise_sum_matches_numpy(dtype, n):
    gen = np.random.RandomState(n)
    for _ in range(50):
        a = gen.rand(n).astype(dtype)
        assert orc.pairwise_sum(a).tobytes() == np.sum(a).tobytes()


def test_play_policy_matches_numpy_for_config_temperatures():
    gen = np.random.RandomState(3)
    for T in (0.0, 0.1, 0.25, 0.5, 1.0):
        for _ in range(50):
            v = gen.randint(0, 200, size=int(gen.choice([2, 10, 82]))).astype(np.int32)
            v[0] += 1
            want = np.asarray(v, dtype=np.int64)
            if T > 0:
                want = np.power(want, max(1.0, min(5.0, 1.0 / T)))
            want = want / np.sum(want)
            assert np.array_equal(bits(want), bits(orc.play_policy(v, T)))


def test_temperature_validation_matches_reference_message():
    with pytest.raises(ValueError, match='Expect `temperature`'):
        orc.play_policy(np.array([1, 2]), 2.0)
    with pytest.raises(ValueError, match='Expect `temperature`'):
        orc.play_policy(np.array([1, 2]), 1)
