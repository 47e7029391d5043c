"""Generate the network fixtures from the UNMODIFIED reference networks.

Run in the build container (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_nets.py

Writes
  tests/golden/ckpt_<name>.npz   the ``network`` state_dict of each checkpoint under
                                 /root/reference/saved_checkpoints (weights only — the
                                 fixtures BASELINE.json's configs 1 and 2 name);
  tests/golden/net_golden.npz    inputs and the reference's own outputs
                                 (``initial_inference`` then a chain of
                                 ``recurrent_inference`` calls) for the three MLP
                                 checkpoints and for seeded random-init ResNets.
and checks, failing loudly otherwise, that
  * oracle/network_oracle.py reproduces the reference modules bit-for-bit;
  * muzero_b200.network's modules have the reference's state_dict keys and shapes,
    and — built under the same ``torch.manual_seed`` — the same initial weights.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')
sys.dont_write_bytecode = True

from muzero import network as ref_net                    # noqa: E402  (the reference)

import muzero_b200.network as my_net                     # noqa: E402
from oracle.network_oracle import OracleNet, randomize_batchnorm   # noqa: E402

CKPT = '/root/reference/saved_checkpoints'
MLPS = {
    'tictactoe': ('TicTacToe_train_steps_35000', dict(input_shape=(9, 3, 3), num_actions=10, num_planes=256,
                                                      value_support_size=1, reward_support_size=1, hidden_dim=64)),
    'cartpole': ('CartPole-v1_train_steps_44800', dict(input_shape=(4, 5), num_actions=2, num_planes=512,
                                                       value_support_size=31, reward_support_size=31, hidden_dim=64)),
    'lunarlander': ('LunarLander-v2_train_steps_58400', dict(input_shape=(4, 9), num_actions=4, num_planes=512,
                                                             value_support_size=31, reward_support_size=31,
                                                             hidden_dim=64)),
}
# seeded random-init conv nets: (kind, ctor kwargs, seed)
CONVS = {
    'board_small': ('board', dict(input_shape=(5, 5, 5), num_actions=26, num_res_blocks=2, num_planes=32), 3),
    'gomoku': ('board', dict(input_shape=(9, 9, 9), num_actions=82, num_res_blocks=8, num_planes=128), 0),
    'atari_small': ('atari', dict(input_shape=(4, 96, 96), num_actions=6, num_res_blocks=2, num_planes=128,
                                  value_support_size=21, reward_support_size=21), 5),
}


def chain(net, obs, actions):
    """initial_inference + len(actions) chained recurrent_inference calls, reference API, batch 1."""
    out = {}
    o = net.initial_inference(torch.from_numpy(obs)[None])
    out['h0'], out['pi0'], out['v0'] = o.hidden_state, o.pi_probs, np.float32(o.value)
    h = o.hidden_state
    hs, rs, vs, pis = [], [], [], []
    for a in actions:
        o = net.recurrent_inference(torch.from_numpy(h)[None], torch.tensor([[int(a)]], dtype=torch.long))
        h = o.hidden_state
        hs.append(h); rs.append(np.float32(o.reward)); vs.append(np.float32(o.value)); pis.append(o.pi_probs)
    out['h'], out['r'], out['v'], out['pi'] = np.stack(hs), np.array(rs), np.array(vs), np.stack(pis)
    return out


def same(a, b, what):
    a, b = np.atleast_1d(np.asarray(a)), np.atleast_1d(np.asarray(b))
    assert a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8)), f'{what}: oracle != reference'


def check_module_parity(ref, mine, what):
    rs, ms = ref.state_dict(), mine.state_dict()
    assert list(rs.keys()) == list(ms.keys()), f'{what}: state_dict keys differ'
    for k in rs:
        assert rs[k].shape == ms[k].shape, f'{what}: shape of {k}'
        assert torch.equal(rs[k], ms[k]), f'{what}: same seed, different initial value of {k}'


def main():
    store = {}
    rs = np.random.RandomState(7)
    for name, (fname, kw) in MLPS.items():
        ck = torch.load(os.path.join(CKPT, fname), map_location='cpu', weights_only=False)
        sd = {k: v.float() for k, v in ck['network'].items()}
        np.savez_compressed(os.path.join(HERE, f'ckpt_{name}.npz'), **{k: v.numpy() for k, v in sd.items()})
        ref = ref_net.MuZeroMLPNet(**kw)
        ref.load_state_dict(sd)
        ref.eval()
        mine = my_net.MuZeroMLPNet(**kw)
        mine.load_state_dict(sd)                      # the reference checkpoint loads unchanged
        torch.manual_seed(11); r2 = ref_net.MuZeroMLPNet(**kw)
        torch.manual_seed(11); m2 = my_net.MuZeroMLPNet(**kw)
        check_module_parity(r2, m2, name)
        orc = OracleNet('mlp', sd, kw['num_actions'], kw['value_support_size'], kw['reward_support_size'])
        for j in range(4):
            obs = rs.standard_normal(kw['input_shape']).astype(np.float32) if name != 'tictactoe' else \
                rs.randint(0, 2, size=kw['input_shape']).astype(np.float32)
            acts = rs.randint(0, kw['num_actions'], size=6)
            g, o = chain(ref, obs, acts), chain(orc, obs, acts)
            for k in g:
                same(g[k], o[k], f'{name}/{j}/{k}')
            store[f'{name}_{j}_obs'], store[f'{name}_{j}_actions'] = obs, acts
            for k in g:
                store[f'{name}_{j}_{k}'] = g[k]
        print(f'{name}: checkpoint exported, oracle == reference, module keys/init == reference')

    for name, (kind, kw, seed) in CONVS.items():
        rcls = ref_net.MuZeroBoardGameNet if kind == 'board' else ref_net.MuZeroAtariNet
        mcls = my_net.MuZeroBoardGameNet if kind == 'board' else my_net.MuZeroAtariNet
        torch.manual_seed(seed); ref = rcls(**kw).eval()
        torch.manual_seed(seed); mine = mcls(**kw).eval()
        check_module_parity(ref, mine, name)
        # non-trivial BatchNorm statistics, deterministic, reproduced by the tests from the same recipe
        randomize_batchnorm(ref, 1000 + seed)
        sd = ref.state_dict()
        orc = OracleNet(kind, sd, kw['num_actions'], kw.get('value_support_size', 1),
                        kw.get('reward_support_size', 1), kw['num_res_blocks'])
        c = kw['input_shape'][0]
        for j in range(2):
            if kind == 'board':
                obs = rs.randint(0, 2, size=kw['input_shape']).astype(np.float32)
            else:
                obs = rs.randint(0, 256, size=kw['input_shape']).astype(np.float32)
                obs[c // 2:] = (rs.randint(0, kw['num_actions'], size=(c - c // 2, 1, 1)) + 1) / kw['num_actions']
            acts = rs.randint(0, kw['num_actions'], size=3)
            gg, o = chain(ref, obs, acts), chain(orc, obs, acts)
            for k in gg:
                same(gg[k], o[k], f'{name}/{j}/{k}')
            store[f'{name}_{j}_obs'], store[f'{name}_{j}_actions'] = obs, acts
            for k in gg:
                store[f'{name}_{j}_{k}'] = gg[k].astype(np.float16) if k in ('h0', 'h') else gg[k]
        print(f'{name}: oracle == reference, module keys/init == reference')
    out = os.path.join(HERE, 'net_golden.npz')
    np.savez_compressed(out, **store)
    print(f'wrote {out} ({os.path.getsize(out)} bytes)')


if __name__ == '__main__':
    main()
