"""Record the reference's own `calc_loss` (pipeline.py:541-612) for a network the hand-written training towers cover
(csrc/train.cu): MuZeroBoardGameNet((9,9,9), 82, 2 blocks, 128 planes), batch 16, K = 5.

Run in the build container:  PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_train_r2.py
Writes tests/golden/train_golden_r2.npz: loss, priorities and, per parameter, the gradient's L2 norm and its first 256
entries (the file stays small; the norm pins the scale, the entries the direction)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_train as base                      # noqa: E402  (installs the gym / snappy stand-ins, imports the reference)

KW = dict(input_shape=(9, 9, 9), num_actions=82, num_res_blocks=2, num_planes=128)
B, T, SEED = 16, 5, 78


def main():
    torch.manual_seed(23); ref = base.ref_net.MuZeroBoardGameNet(**KW)
    torch.manual_seed(23); mine = base.my_net.MuZeroBoardGameNet(**KW)
    ref.train(); mine.train()
    for (k, a), (_, b) in zip(ref.state_dict().items(), mine.state_dict().items()):
        assert torch.equal(a, b), k
    tr, w = base.synthetic_transitions(mine, B, T, seed=SEED)
    rtr = base.RefTransition(state=tr.state, action=tr.action, pi_prob=tr.pi_prob, value=tr.value, reward=tr.reward)
    loss, pri = base.ref_pipe.calc_loss(ref, 'cpu', rtr, torch.from_numpy(w))
    loss.backward()
    store = {'loss': np.float64(loss.item()), 'priorities': pri}
    for k, p in ref.named_parameters():
        g = p.grad.detach().reshape(-1)
        store[f'grad_norm_{k}'] = np.float64(g.double().norm().item())
        store[f'grad_head_{k}'] = g[:256].numpy().copy()
    for k, b in ref.named_buffers():
        if 'running' in k:
            store[f'buffer_{k}'] = b.detach().numpy().copy()
    out = os.path.join(HERE, 'train_golden_r2.npz')
    np.savez_compressed(out, **store)
    print('loss', loss.item(), 'wrote', out, os.path.getsize(out))


if __name__ == '__main__':
    main()
