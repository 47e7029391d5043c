"""Round-2 network fixtures from the UNMODIFIED reference networks: the shapes the reference BUILDS that round 1
did not record (run in the build container, needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_nets_r2.py

Writes tests/golden/net_golden_r2.npz (the existing net_golden.npz is left byte-for-byte as it was) with inputs and
the reference's own outputs -- ``initial_inference`` then a chain of ``recurrent_inference`` calls, batch 1 -- for
  atari_c4       MuZeroAtariNet((16, 96, 96), 18, 8, 128, 61, 61)   BASELINE.json configs[3] / config.py:204-233
  ttt_resnet     MuZeroBoardGameNet((9, 3, 3), 10, 2, 16)           config.py:126-127 (use_mlp_net=False)
  board_256x16   MuZeroBoardGameNet((9, 9, 9), 82, 16, 256)         class defaults, network.py:543-549
  atari_default  MuZeroAtariNet((4, 96, 96), 6, 16, 256, 601, 601)  class defaults, network.py:504-512
and checks, failing loudly otherwise, that oracle/network_oracle.py reproduces the reference bit-for-bit on them and
that muzero_b200.network builds the same state_dict (keys, shapes, same-seed initial values).
Atari frames are stored as uint8 + the per-plane action fractions (the float32 observation is rebuilt by the test).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, '/root/reference')
sys.dont_write_bytecode = True

from muzero import network as ref_net                    # noqa: E402  (the reference)

import muzero_b200.network as my_net                     # noqa: E402
from make_golden_nets import chain, check_module_parity, same      # noqa: E402
from oracle.network_oracle import OracleNet, randomize_batchnorm   # noqa: E402

CONVS = {
    'atari_c4': ('atari', dict(input_shape=(16, 96, 96), num_actions=18, num_res_blocks=8, num_planes=128,
                               value_support_size=61, reward_support_size=61), 0),
    'ttt_resnet': ('board', dict(input_shape=(9, 3, 3), num_actions=10, num_res_blocks=2, num_planes=16), 7),
    'board_256x16': ('board', dict(input_shape=(9, 9, 9), num_actions=82, num_res_blocks=16, num_planes=256), 8),
    'atari_default': ('atari', dict(input_shape=(4, 96, 96), num_actions=6, num_res_blocks=16, num_planes=256,
                                    value_support_size=601, reward_support_size=601), 9),
}


def atari_obs(frames_u8: np.ndarray, fracs: np.ndarray) -> np.ndarray:
    """[c/2] uint8 frames + [c - c/2] action fractions -> float32 [c, 96, 96] (gym_env.py stacks frames then
    broadcast action planes)."""
    planes = np.broadcast_to(fracs.astype(np.float32)[:, None, None], (len(fracs),) + frames_u8.shape[1:])
    return np.concatenate([frames_u8.astype(np.float32), planes], 0)


def main():
    store = {}
    rs = np.random.RandomState(8)
    for name, (kind, kw, seed) in CONVS.items():
        rcls = ref_net.MuZeroBoardGameNet if kind == 'board' else ref_net.MuZeroAtariNet
        mcls = my_net.MuZeroBoardGameNet if kind == 'board' else my_net.MuZeroAtariNet
        torch.manual_seed(seed); ref = rcls(**kw).eval()
        torch.manual_seed(seed); mine = mcls(**kw).eval()
        check_module_parity(ref, mine, name)
        randomize_batchnorm(ref, 1000 + seed)
        sd = ref.state_dict()
        orc = OracleNet(kind, sd, kw['num_actions'], kw.get('value_support_size', 1),
                        kw.get('reward_support_size', 1), kw['num_res_blocks'])
        c = kw['input_shape'][0]
        for j in range(2):
            if kind == 'board':
                obs = rs.randint(0, 2, size=kw['input_shape']).astype(np.float32)
                store[f'{name}_{j}_obs'] = obs.astype(np.int8)
            else:
                frames = rs.randint(0, 256, size=(c // 2,) + kw['input_shape'][1:]).astype(np.uint8)
                fracs = ((rs.randint(0, kw['num_actions'], size=c - c // 2) + 1) / kw['num_actions']).astype(np.float32)
                obs = atari_obs(frames, fracs)
                store[f'{name}_{j}_frames'], store[f'{name}_{j}_fracs'] = frames, fracs
            acts = rs.randint(0, kw['num_actions'], size=3)
            gg, o = chain(ref, obs, acts), chain(orc, obs, acts)
            for k in gg:
                same(gg[k], o[k], f'{name}/{j}/{k}')
            store[f'{name}_{j}_actions'] = acts
            for k in gg:
                store[f'{name}_{j}_{k}'] = gg[k].astype(np.float16) if k in ('h0', 'h') else gg[k]
        print(f'{name}: oracle == reference, module keys/init == reference', flush=True)
    out = os.path.join(HERE, 'net_golden_r2.npz')
    np.savez_compressed(out, **store)
    print(f'wrote {out} ({os.path.getsize(out)} bytes)')


if __name__ == '__main__':
    main()
