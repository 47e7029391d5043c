"""TEST INFRASTRUCTURE ONLY — numpy/CPython restatement of the reference search.

Restates ``muzero/mcts.py`` of michaelnny/muzero as a struct-of-arrays tree
with explicit arithmetic, explicit MT19937 draws and an explicit pairwise sum,
i.e. in the shape the CUDA kernels use.  Each function cites the reference
lines it follows.  Checked bit-for-bit against the imported reference by
``tests/golden/make_golden.py`` (build container) and against the committed
fixtures by ``tests/test_oracle_golden.py`` (anywhere).

The parity target is the reference executed under numpy >= 2 (NEP-50 scalar
promotion) — see ``child_U`` below.
"""
from __future__ import annotations

import math
from typing import NamedTuple, Optional, Sequence

import numpy as np

f32 = np.float32


# --------------------------------------------------------------------------
# Legacy numpy MT19937 stream (np.random.* / np.random.RandomState)
# --------------------------------------------------------------------------
class MT19937:
    """The generator behind ``np.random.seed/choice/dirichlet`` restated.

    numpy is a third-party dependency of the reference (pinned numpy==1.21.6 in
    requirements.txt:21, executed here with numpy 2.x); the legacy stream is the
    published MT19937 of Matsumoto & Nishimura plus numpy's
    ``legacy-distributions.c`` samplers.  Call sites in the reference:
    mcts.py:124 (tie-break ``choice``), mcts.py:245 (``dirichlet``),
    mcts.py:404 (``choice(p=...)``).
    """

    N, M = 624, 397

    def __init__(self, key: np.ndarray, pos: int):
        self.key = np.array(key, dtype=np.uint32).copy()
        assert self.key.shape == (624,)
        self.pos = int(pos)

    # -- construction / hand-back -----------------------------------------
    @classmethod
    def from_seed(cls, seed: int) -> "MT19937":
        """``init_genrand`` (what ``np.random.seed(int)`` does)."""
        key = np.empty(624, dtype=np.uint32)
        s = seed & 0xFFFFFFFF
        for i in range(624):
            key[i] = s
            s = (1812433253 * (s ^ (s >> 30)) + i + 1) & 0xFFFFFFFF
        return cls(key, 624)

    @classmethod
    def from_numpy(cls, rs=None) -> "MT19937":
        st = (np.random if rs is None else rs).get_state()
        assert st[0] == 'MT19937'
        return cls(st[1], st[2])

    def to_numpy(self, rs=None) -> None:
        """Write the advanced state back so the numpy stream continues."""
        tgt = np.random if rs is None else rs
        old = tgt.get_state()
        tgt.set_state(('MT19937', self.key.copy(), self.pos, old[3], old[4]))

    # -- raw draws ----------------------------------------------------------
    def _twist(self) -> None:
        k = self.key.astype(np.uint64)
        UP, LO, A, Z = np.uint64(0x80000000), np.uint64(0x7FFFFFFF), np.uint64(0x9908B0DF), np.uint64(0)
        one = np.uint64(1)
        N, M = self.N, self.M
        # three dependency-free phases (the CUDA twist uses the same split)
        for lo, hi in ((0, N - M), (N - M, 2 * (N - M)), (2 * (N - M), N - 1)):
            i = np.arange(lo, hi)
            y = (k[i] & UP) | (k[i + 1] & LO)
            src = np.where(i < N - M, i + M, i + M - N)
            k[i] = k[src] ^ (y >> one) ^ np.where((y & one) != Z, A, Z)
        y = (k[N - 1] & UP) | (k[0] & LO)
        k[N - 1] = k[M - 1] ^ (y >> one) ^ (A if (y & one) else Z)
        self.key = k.astype(np.uint32)
        self.pos = 0

    def next_u32(self) -> int:
        if self.pos >= 624:
            self._twist()
        y = int(self.key[self.pos])
        self.pos += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF

    def next_double(self) -> float:
        """``random_sample`` / ``legacy_double``: 53 bits from two draws."""
        a = self.next_u32() >> 5
        b = self.next_u32() >> 6
        return (a * 67108864.0 + b) / 9007199254740992.0

    # -- numpy legacy distributions used by the reference ------------------
    def bounded(self, k: int) -> int:
        """``randint(0, k)`` as used by ``choice(array_of_k)`` (mcts.py:124):
        no draw when k == 1, else masked rejection on 32-bit draws."""
        rng = k - 1
        if rng == 0:
            return 0
        mask = rng
        mask |= mask >> 1
        mask |= mask >> 2
        mask |= mask >> 4
        mask |= mask >> 8
        mask |= mask >> 16
        while True:
            v = self.next_u32() & mask
            if v <= rng:
                return v

    def standard_exponential(self) -> float:
        return -math.log(1.0 - self.next_double())

    def standard_gamma(self, shape: float) -> float:
        """``legacy_standard_gamma`` for shape <= 1 (the reference restricts
        alpha to [0, 1], mcts.py:241-242)."""
        if shape == 1.0:
            return self.standard_exponential()
        if shape == 0.0:
            return 0.0
        if shape > 1.0:
            raise NotImplementedError('alpha > 1 is rejected by the reference (mcts.py:241)')
        while True:
            u = self.next_double()
            v = self.standard_exponential()
            if u <= 1.0 - shape:
                x = math.pow(u, 1.0 / shape)
                if x <= v:
                    return x
            else:
                y = -math.log((1.0 - u) / shape)
                x = math.pow(1.0 - shape + shape * y, 1.0 / shape)
                if x <= (v + y):
                    return x

    def dirichlet(self, alphas: np.ndarray) -> np.ndarray:
        """``RandomState.dirichlet`` (legacy path): gamma draws / their sum."""
        k = len(alphas)
        val = np.empty(k, dtype=np.float64)
        acc = 0.0
        for j in range(k):
            val[j] = self.standard_gamma(float(alphas[j]))
            acc += val[j]
        inv = 1.0 / acc
        for j in range(k):
            val[j] = val[j] * inv
        return val

    def choice_p(self, p: np.ndarray) -> int:
        """``choice(A, p=p)`` for size=None (mcts.py:404): one uniform double,
        ``cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(cdf, u, 'right')``."""
        cdf = np.empty(len(p), dtype=np.float64)
        acc = 0.0
        for i in range(len(p)):
            acc = acc + float(p[i])
            cdf[i] = acc
        last = cdf[-1]
        u = self.next_double()
        idx = 0
        for i in range(len(p)):           # count of entries <= u == insertion point 'right'
            if cdf[i] / last <= u:
                idx = i + 1
            else:
                break
        return idx


# --------------------------------------------------------------------------
# numpy pairwise summation (np.sum over a contiguous 1-D float array)
# --------------------------------------------------------------------------
def pairwise_sum(a: np.ndarray):
    """``np.sum`` of a contiguous 1-D float32/float64 array, restated
    (numpy ``pairwise_sum_@TYPE@``; used by mcts.py:296 and mcts.py:279)."""
    t = a.dtype.type
    n = len(a)
    if n < 8:
        # numpy starts from -0.0 so that sum([-0.0]) keeps its sign
        res = t(-0.0)
        for i in range(n):
            res = t(res + a[i])
        return res
    if n <= 128:
        r = [t(a[i]) for i in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = t(r[j] + a[i + j])
            i += 8
        res = t(t(t(r[0] + r[1]) + t(r[2] + r[3])) + t(t(r[4] + r[5]) + t(r[6] + r[7])))
        while i < n:
            res = t(res + a[i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return t(pairwise_sum(a[:n2]) + pairwise_sum(a[n2:]))


# --------------------------------------------------------------------------
# Root prior preparation
# --------------------------------------------------------------------------
def mix_dirichlet(prior: np.ndarray, noise: np.ndarray, eps: float) -> np.ndarray:
    """mcts.py:244-246.  ``(1 - eps) * prob`` stays float32 for a float32
    prior (python scalar is weak under NEP-50); ``eps * noise`` is float64; the
    sum promotes to float64."""
    if prior.dtype == np.float32:
        left = (f32(1 - eps) * prior).astype(np.float64)
    else:
        left = (1 - eps) * prior
    return left + eps * noise.astype(np.float64)


def mask_and_renormalise(mask: np.ndarray, prob: np.ndarray) -> np.ndarray:
    """mcts.py:283-299: zero the illegal entries, divide by the pairwise sum
    when it is positive; dtype (float32 / float64) is preserved."""
    assert mask.shape == prob.shape
    t = prob.dtype.type
    out = np.where(mask, prob, t(0.0)).astype(prob.dtype)
    s = pairwise_sum(out)
    if s > 0:
        out = (out / s).astype(prob.dtype)
    return out


def pb_c_table(num_simulations: int, pb_c_base: float, pb_c_init: float) -> np.ndarray:
    """``(log((N + base + 1) / base) + init) * sqrt(N)`` of mcts.py:193-195 for
    every possible parent visit count, with CPython ``math`` (float64)."""
    return np.array(
        [(math.log((n + pb_c_base + 1) / pb_c_base) + pb_c_init) * math.sqrt(n)
         for n in range(num_simulations + 2)], dtype=np.float64)


# --------------------------------------------------------------------------
# The tree
# --------------------------------------------------------------------------
class SearchTrace(NamedTuple):
    """Everything a parity test compares, struct-of-arrays."""
    num_nodes: int
    N: np.ndarray          # int32  [nodes]
    W: np.ndarray          # float64[nodes]
    R: np.ndarray          # float64[nodes]  (float32-representable)
    parent: np.ndarray     # int32  [nodes]
    move: np.ndarray       # int32  [nodes]
    children: np.ndarray   # int32  [nodes, A]  (-1 = never visited)
    depth: np.ndarray      # int32  [nodes]
    prior: np.ndarray      # float32 or float64 [A]  (the ONE prior every node uses)
    minmax: tuple          # (min, max) after the search
    sel_parent: np.ndarray  # int32 [S] node each simulation expanded from
    sel_action: np.ndarray  # int32 [S]
    sel_depth: np.ndarray   # int32 [S] depth of the new leaf
    tie_draws: int          # number of 32-bit draws consumed by tie-breaks


class OracleTree:
    """SoA restatement of ``Node`` + ``MinMaxStats`` (mcts.py:33-217)."""

    def __init__(self, num_actions, num_simulations, prior, discount, is_board_game,
                 known_bounds, pb_c_base, pb_c_init, current_player, opponent_player, rng: MT19937):
        A, S = num_actions, num_simulations
        self.A, self.S = A, S
        self.prior = prior
        self.f32path = prior.dtype == np.float32
        self.discount = float(discount)
        self.board = bool(is_board_game)
        self.dp = self.discount * (-1.0 if self.board else 1.0)      # mcts.py:169-174
        # mcts.py:36-38 (KnownBounds holds python ints for the board games; int/float
        # comparisons and arithmetic below are exact either way)
        self.hi = float(known_bounds.max) if known_bounds else -math.inf
        self.lo = float(known_bounds.min) if known_bounds else math.inf
        self.T = pb_c_table(S, pb_c_base, pb_c_init)
        self.cur, self.opp = current_player, opponent_player
        self.rng = rng
        n = S + 1
        self.Nv = np.zeros(n, dtype=np.int32)
        self.W = np.zeros(n, dtype=np.float64)
        self.R = np.zeros(n, dtype=np.float64)
        self.PL = np.zeros(n, dtype=np.int64)
        self.PAR = np.full(n, -1, dtype=np.int32)
        self.MOVE = np.full(n, -1, dtype=np.int32)
        self.DEPTH = np.zeros(n, dtype=np.int32)
        self.CH = np.full((n, A), -1, dtype=np.int32)
        self.hidden = [None] * n
        self.count = 0
        self.tie_draws = 0
        self.sel_parent, self.sel_action, self.sel_depth = [], [], []

    # mcts.py:75-102 — unvisited children are N=0, W=0, reward=0, so a child
    # slot is only materialised on its first visit (observationally identical).
    def set_root(self, hidden, reward):
        self.count = 1
        self.R[0] = reward
        self.PL[0] = self.cur
        self.hidden[0] = hidden

    def ucb_scores(self, n: int) -> np.ndarray:
        """``child_Q + child_U`` (mcts.py:121,159-200) in float32."""
        A = self.A
        t = self.T[self.Nv[n]]
        ucb = np.empty(A, dtype=np.float32)
        for a in range(A):
            c = self.CH[n, a]
            cn = int(self.Nv[c]) if c >= 0 else 0
            y = t / (cn + 1)
            if self.f32path:
                u = f32(self.prior[a]) * f32(y)               # NEP-50: float32 x weak python float
            else:
                u = f32(float(self.prior[a]) * y)             # float64 product, then array cast
            if cn > 0:
                v = self.R[c] + self.dp * (self.W[c] / cn)
                if self.hi > self.lo:                         # mcts.py:44-48
                    v = (v - self.lo) / (self.hi - self.lo)
                q = f32(v)
            else:
                q = f32(0.0)
            ucb[a] = q + u                                    # float32 + float32
        return ucb

    def select(self):
        """mcts.py:372-379."""
        n, cp, op = 0, self.cur, self.opp
        depth = 0
        while True:
            ucb = self.ucb_scores(n)
            ties = np.where(ucb == ucb.max())[0]
            k = len(ties)
            before = self.rng.pos
            a = int(ties[self.rng.bounded(k)])                # mcts.py:124
            if k > 1:
                self.tie_draws += 1  # counts tie events; raw draw count is in the stream itself
            cp, op = op, cp
            depth += 1
            if self.CH[n, a] < 0:
                return n, a, cp, depth
            n = int(self.CH[n, a])

    def expand_backup(self, n, a, leaf_player, depth, hidden, reward, value):
        """mcts.py:386 (expand) and mcts.py:129-157 (backup)."""
        c = self.count
        self.count += 1
        self.CH[n, a] = c
        self.PAR[c], self.MOVE[c], self.DEPTH[c] = n, a, depth
        self.hidden[c] = hidden
        self.R[c] = reward
        self.PL[c] = leaf_player
        self.sel_parent.append(n); self.sel_action.append(a); self.sel_depth.append(depth)
        x, pid = c, leaf_player
        value = float(value)
        while x >= 0:
            self.W[x] += value if self.PL[x] == pid else -value
            self.Nv[x] += 1
            q = self.W[x] / int(self.Nv[x])
            mm = self.R[x] + self.discount * (-q if self.board else q)
            self.hi = max(self.hi, mm)
            self.lo = min(self.lo, mm)
            if self.board and self.PL[x] == pid:
                value = -self.R[x] + self.discount * value
            else:
                value = self.R[x] + self.discount * value
            x = int(self.PAR[x])

    def root_visits(self) -> np.ndarray:
        out = np.zeros(self.A, dtype=np.int32)
        for a in range(self.A):
            c = self.CH[0, a]
            out[a] = self.Nv[c] if c >= 0 else 0
        return out

    def trace(self) -> SearchTrace:
        k = self.count
        return SearchTrace(k, self.Nv[:k].copy(), self.W[:k].copy(), self.R[:k].copy(), self.PAR[:k].copy(),
                           self.MOVE[:k].copy(), self.CH[:k].copy(), self.DEPTH[:k].copy(), self.prior.copy(),
                           (self.lo, self.hi), np.array(self.sel_parent, np.int32),
                           np.array(self.sel_action, np.int32), np.array(self.sel_depth, np.int32), self.tie_draws)


def play_policy(visits: np.ndarray, temperature: float) -> np.ndarray:
    """mcts.py:250-280.  int64 counts, optional ``** clamp(1/T, 1, 5)`` in
    float64, divided by the pairwise sum."""
    if not isinstance(temperature, float) or not 0.0 <= temperature <= 1.0:
        raise ValueError(f'Expect `temperature` to be float type in the range [0.0, 1.0], got {temperature}')
    v = np.asarray(visits, dtype=np.int64)
    if temperature > 0.0:
        e = max(1.0, min(5.0, 1.0 / temperature))
        if e == int(e):
            # exact integer power, correctly rounded (every temperature the reference's
            # configs use -- 1.0, 0.5, 0.25, 0.1, config.py:236-267 -- lands here)
            v = np.array([float(int(x) ** int(e)) for x in v], dtype=np.float64)
        else:
            # np.power dispatches to SVML or libm by CPU and is only good to an ulp or two:
            # the reference is not reproducible across machines here, parity is <= 4 ulp.
            v = np.array([math.pow(float(x), e) for x in v], dtype=np.float64)
        s = pairwise_sum(v)
        return v / s
    # integer path: np.sum of int64 is exact, the division is float64
    s = int(v.sum())
    with np.errstate(invalid='ignore', divide='ignore'):
        return v / np.int64(s)


def uct_search(state, network, device, config, temperature, actions_mask, current_player, opponent_player,
               deterministic: bool = False, rng=None, noise: Optional[np.ndarray] = None,
               return_trace: bool = False):
    """Restatement of ``uct_search`` (mcts.py:302-407), same signature plus:

    rng   : ``np.random.RandomState`` to draw from (default: the global
            ``np.random`` stream, exactly like the reference); its state is
            advanced exactly as the reference would advance it.
    noise : optional pre-drawn Dirichlet sample (float64[A]) used INSTEAD of
            drawing one — the "identical noise" injection point of the parity
            harness.
    """
    import torch
    if config.is_board_game:
        assert config.discount == 1.0
    mt = MT19937.from_numpy(rng)

    st = torch.from_numpy(np.asarray(state)).to(device=device, dtype=torch.float32)
    out0 = network.initial_inference(st[None, ...])
    prior = np.asarray(out0.pi_probs)
    A = prior.shape[0]

    if not deterministic and config.root_dirichlet_alpha > 0.0 and config.root_exploration_eps > 0.0:
        eps, alpha = config.root_exploration_eps, config.root_dirichlet_alpha
        if not isinstance(eps, float) or not 0.0 <= eps <= 1.0:
            raise ValueError(f'Expect `eps` to be a float in the range [0.0, 1.0], got {eps}')
        if not isinstance(alpha, float) or not 0.0 <= alpha <= 1.0:
            raise ValueError(f'Expect `alpha` to be a float in the range [0.0, 1.0], got {alpha}')
        if noise is None:
            noise = mt.dirichlet((np.ones_like(prior) * alpha).astype(np.float64))
        prior = mix_dirichlet(prior, np.asarray(noise), eps)
    if actions_mask is not None:
        prior = mask_and_renormalise(np.asarray(actions_mask), prior)

    tree = OracleTree(A, config.num_simulations, prior, config.discount, config.is_board_game,
                      config.known_bounds, config.pb_c_base, config.pb_c_init,
                      current_player, opponent_player, mt)
    tree.set_root(out0.hidden_state, out0.reward)

    for _ in range(config.num_simulations):
        n, a, leaf_player, depth = tree.select()
        h = torch.from_numpy(np.asarray(tree.hidden[n])).to(device=device, dtype=torch.float32)
        act = torch.tensor([a], dtype=torch.long, device=device)
        o = network.recurrent_inference(h[None, ...], act[None, ...])
        tree.expand_backup(n, a, leaf_player, depth, o.hidden_state, o.reward, o.value)

    visits = tree.root_visits()
    if actions_mask is not None:
        visits = np.where(np.asarray(actions_mask), visits, 0)
    pi = play_policy(visits, temperature)
    if deterministic:
        action = int(np.argmax(visits))
    else:
        if np.isnan(pi).any():
            raise ValueError('probabilities contain NaN')
        action = mt.choice_p(pi)
    mt.to_numpy(rng)
    root_q = float(tree.W[0] / int(tree.Nv[0])) if tree.Nv[0] > 0 else 0.0
    if return_trace:
        return action, pi, root_q, tree.trace()
    return action, pi, root_q
