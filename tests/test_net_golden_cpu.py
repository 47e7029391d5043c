"""CPU pin of the network oracle: oracle/network_oracle.py against the reference's own recorded outputs
(tests/golden/net_golden.npz and net_golden_r2.npz, written by make_golden_nets*.py from the unmodified reference).
The generators assert bit equality inside one process; across machines torch's CPU convolutions may pick another
algorithm, so the bar here is 1e-4 relative (hidden states are stored as float16: 1e-3 absolute on [0, 1])."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle.network_oracle import OracleNet, randomize_batchnorm

R1 = {
    'board_small': ('board', dict(input_shape=(5, 5, 5), num_actions=26, num_res_blocks=2, num_planes=32), 3),
    'gomoku': ('board', dict(input_shape=(9, 9, 9), num_actions=82, num_res_blocks=8, num_planes=128), 0),
    'atari_small': ('atari', dict(input_shape=(4, 96, 96), num_actions=6, num_res_blocks=2, num_planes=128,
                                  value_support_size=21, reward_support_size=21), 5),
}
R2 = {
    'atari_c4': ('atari', dict(input_shape=(16, 96, 96), num_actions=18, num_res_blocks=8, num_planes=128,
                               value_support_size=61, reward_support_size=61), 0),
    'ttt_resnet': ('board', dict(input_shape=(9, 3, 3), num_actions=10, num_res_blocks=2, num_planes=16), 7),
    'board_256x16': ('board', dict(input_shape=(9, 9, 9), num_actions=82, num_res_blocks=16, num_planes=256), 8),
    'atari_default': ('atari', dict(input_shape=(4, 96, 96), num_actions=6, num_res_blocks=16, num_planes=256,
                                    value_support_size=601, reward_support_size=601), 9),
}


def golden_obs(z, name, j):
    """Observation j of a recorded case (round-2 Atari cases store uint8 frames + action fractions)."""
    if f'{name}_{j}_obs' in z:
        return z[f'{name}_{j}_obs'].astype(np.float32)
    frames, fracs = z[f'{name}_{j}_frames'], z[f'{name}_{j}_fracs']
    planes = np.broadcast_to(fracs.astype(np.float32)[:, None, None], (len(fracs),) + frames.shape[1:])
    return np.concatenate([frames.astype(np.float32), planes], 0)


def build_pair(kind, kw, seed):
    """(muzero_b200 module on the CPU, OracleNet) with the weights the recording was made with: same torch seed as the
    reference build (the generator checks the initial values are identical) + the deterministic BatchNorm recipe."""
    import muzero_b200 as mz
    torch.manual_seed(seed)
    cls = mz.MuZeroBoardGameNet if kind == 'board' else mz.MuZeroAtariNet
    net = cls(**kw).eval()
    randomize_batchnorm(net, 1000 + seed)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    onet = OracleNet(kind, sd, kw['num_actions'], kw.get('value_support_size', 1), kw.get('reward_support_size', 1),
                     kw['num_res_blocks'])
    return net, onet


@pytest.mark.parametrize('fname,name', [('net_golden.npz', n) for n in R1] + [('net_golden_r2.npz', n) for n in R2])
def test_network_oracle_reproduces_the_reference_recordings(fname, name):
    kind, kw, seed = (R1 if name in R1 else R2)[name]
    z = np.load(os.path.join(GOLDEN, fname))
    _, onet = build_pair(kind, kw, seed)
    for j in range(2):
        obs = golden_obs(z, name, j)
        o = onet.initial_inference(torch.from_numpy(obs)[None])
        np.testing.assert_allclose(o.hidden_state, z[f'{name}_{j}_h0'].astype(np.float32), atol=1e-3)
        np.testing.assert_allclose(o.pi_probs, z[f'{name}_{j}_pi0'], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(o.value, z[f'{name}_{j}_v0'], rtol=1e-4, atol=1e-5)
        h = o.hidden_state
        for i, a in enumerate(z[f'{name}_{j}_actions']):
            o = onet.recurrent_inference(torch.from_numpy(h)[None], torch.tensor([[int(a)]]))
            np.testing.assert_allclose(o.hidden_state, z[f'{name}_{j}_h'][i].astype(np.float32), atol=1e-3)
            np.testing.assert_allclose(o.reward, z[f'{name}_{j}_r'][i], rtol=1e-4, atol=1e-5)
            np.testing.assert_allclose(o.value, z[f'{name}_{j}_v'][i], rtol=1e-4, atol=1e-5)
            np.testing.assert_allclose(o.pi_probs, z[f'{name}_{j}_pi'][i], rtol=1e-4, atol=1e-6)
            h = o.hidden_state
