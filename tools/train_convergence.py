"""End-to-end sanity of the training kernels beyond single-step parity: the same learner (DataParallelLearner: K=5
unroll, Adam, CUDA graph) fits a fixed set of synthetic replay batches, once with the towers on csrc/train.cu and once
through PyTorch autograd + cuDNN (TF32), from the same initial weights.  Prints both loss curves.
usage: python tools/train_convergence.py [steps] [blocks]"""
import copy
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import muzero_b200 as mz
from muzero_b200.training import DataParallelLearner, synthetic_transitions

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
blocks = int(sys.argv[2]) if len(sys.argv) > 2 else 8
torch.manual_seed(0)
base = mz.MuZeroBoardGameNet((9, 9, 9), 82, blocks, 128).cuda()
cfg = mz.config.make_gomoku_config(num_training_steps=10 * steps, batch_size=128)
batches = [synthetic_transitions(base, 128, 5, seed=100 + k) for k in range(4)]
curves = {}
for name, native in (('kernels', '1'), ('autograd', '0')):
    os.environ['MZ_TRAIN_NATIVE'] = native
    net = copy.deepcopy(base)
    learner = DataParallelLearner(net, cfg, 'cuda')
    assert learner.native_towers == (native == '1')
    losses = []
    for it in range(steps):
        tr, w = batches[it % len(batches)]
        loss, _ = learner.step(tr, w)
        losses.append(loss)
    curves[name] = losses
    del learner, net
    torch.cuda.empty_cache()
os.environ.pop('MZ_TRAIN_NATIVE')
print('step   kernels  autograd   (mean loss over the 4 fixed batches, K=5 unroll, batch 128, %d blocks)' % blocks)
for a in range(0, steps, max(4, steps // 25 // 4 * 4)):
    b = min(steps, a + 4)
    print('%4d  %8.4f  %8.4f' % (a, np.mean(curves['kernels'][a:b]), np.mean(curves['autograd'][a:b])))
print('last  %8.4f  %8.4f' % (np.mean(curves['kernels'][-4:]), np.mean(curves['autograd'][-4:])))
