"""The K=5 training step of BASELINE configs[4] for ncu / timing. usage: python tools/train_step_target.py [steps] [graph 0|1] [blocks]
Prints the mean step time (CUDA events) over the last `steps` iterations."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import muzero_b200 as mz  # noqa: E402
from muzero_b200.training import DataParallelLearner, synthetic_transitions  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
graph = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
blocks = int(sys.argv[3]) if len(sys.argv) > 3 else 8
torch.manual_seed(0)
net = mz.MuZeroBoardGameNet((9, 9, 9), 82, blocks, 128).cuda()
cfg = mz.config.make_gomoku_config(num_training_steps=100, batch_size=128)
learner = DataParallelLearner(net, cfg, 'cuda', use_graph=graph)
tr, w = synthetic_transitions(net, 128, 5, seed=1)
for _ in range(5 if graph else 2):
    loss, _ = learner.step(tr, w)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss, _ = learner.step(tr, w)
e1.record()
torch.cuda.synchronize()
print('native' if learner.native_towers else 'autograd', 'graph' if learner._graph is not None else 'eager',
      'ms/step %.3f' % (e0.elapsed_time(e1) / steps), 'loss %.4f' % loss)
