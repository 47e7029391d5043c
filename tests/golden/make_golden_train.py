"""Record the reference's own `calc_loss` (pipeline.py:541-612) on fixed synthetic replay batches.

Run in the build container:  PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_train.py
`muzero.pipeline` imports gym / snappy, which are not installed: test-side stand-ins are injected into
sys.modules before the import (the reference files are untouched).  Writes tests/golden/train_golden.npz with, per
case, the loss, the priorities and per-parameter gradient digests (sum, |.|-sum, first 8 entries).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')
sys.dont_write_bytecode = True

gym = types.ModuleType('gym')
gym.Env = object
gym.Wrapper = object
gym.ObservationWrapper = object
gym.RewardWrapper = object
gym.spaces = types.ModuleType('gym.spaces')
gym.spaces.Box = gym.spaces.Discrete = object
gym.utils = types.ModuleType('gym.utils')
gym.utils.seeding = types.ModuleType('gym.utils.seeding')
for name, mod in (('gym', gym), ('gym.spaces', gym.spaces), ('gym.utils', gym.utils),
                  ('gym.utils.seeding', gym.utils.seeding), ('snappy', types.ModuleType('snappy'))):
    sys.modules.setdefault(name, mod)
if not hasattr(np, 'bool8'):
    np.bool8 = np.bool_

from muzero import network as ref_net                      # noqa: E402
from muzero import pipeline as ref_pipe                    # noqa: E402
from muzero.replay import Transition as RefTransition      # noqa: E402

import muzero_b200.network as my_net                       # noqa: E402
from muzero_b200.training import calc_loss, synthetic_transitions   # noqa: E402

CASES = {
    'tictactoe_mlp': ('mlp', dict(input_shape=(9, 3, 3), num_actions=10, num_planes=256, value_support_size=1,
                                  reward_support_size=1, hidden_dim=64), 'ckpt_tictactoe.npz', 16),
    'cartpole_mlp': ('mlp', dict(input_shape=(4, 5), num_actions=2, num_planes=512, value_support_size=31,
                                 reward_support_size=31, hidden_dim=64), 'ckpt_cartpole.npz', 16),
    'board_small': ('board', dict(input_shape=(5, 5, 5), num_actions=26, num_res_blocks=2, num_planes=32), None, 12),
}


def digest(net):
    out = {}
    for k, p in net.named_parameters():
        g = p.grad.detach().reshape(-1).double()
        out[k] = np.concatenate([[g.sum().item(), g.abs().sum().item()], g[:8].numpy()])
    return out


def main():
    store = {}
    for name, (kind, kw, ckpt, B) in CASES.items():
        rcls = ref_net.MuZeroMLPNet if kind == 'mlp' else ref_net.MuZeroBoardGameNet
        mcls = my_net.MuZeroMLPNet if kind == 'mlp' else my_net.MuZeroBoardGameNet
        torch.manual_seed(21); ref = rcls(**kw)
        torch.manual_seed(21); mine = mcls(**kw)
        if ckpt:
            sd = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(HERE, ckpt)).items()}
            ref.load_state_dict(sd); mine.load_state_dict(sd)
        ref.train(); mine.train()
        tr, w = synthetic_transitions(mine, B, 5, seed=77)
        rtr = RefTransition(state=tr.state, action=tr.action, pi_prob=tr.pi_prob, value=tr.value, reward=tr.reward)
        loss_r, pri_r = ref_pipe.calc_loss(ref, 'cpu', rtr, torch.from_numpy(w))
        loss_r.backward()
        loss_m, pri_m = calc_loss(mine, 'cpu', tr, torch.from_numpy(w))
        loss_m.backward()
        dr, dm = digest(ref), digest(mine)
        assert abs(loss_r.item() - loss_m.item()) <= 1e-6 * max(1, abs(loss_r.item())), (name, loss_r.item(), loss_m.item())
        assert np.allclose(pri_r, pri_m, rtol=1e-6, atol=1e-7), name
        for k in dr:
            assert np.allclose(dr[k], dm[k], rtol=1e-5, atol=1e-7), (name, k, dr[k], dm[k])
        store[f'{name}_loss'] = np.float64(loss_r.item())
        store[f'{name}_priorities'] = pri_r
        for k, v in dr.items():
            store[f'{name}_grad_{k}'] = v
        print(f'{name}: loss {loss_r.item():.6f}; muzero_b200.training.calc_loss == reference (loss, priorities, {len(dr)} grads)')
    out = os.path.join(HERE, 'train_golden.npz')
    np.savez_compressed(out, **store)
    print('wrote', out, os.path.getsize(out))


if __name__ == '__main__':
    main()
