"""Steady-state time of one tower's forward / backward launch chain on the training kernels (CUDA events, warm caches).
usage: python tools/train_tower_time.py [blocks] [batch] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import muzero_b200 as mz
from muzero_b200 import train_engine

blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 8
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
torch.manual_seed(0)
net = mz.MuZeroBoardGameNet((9, 9, 9), 82, blocks, 128).cuda().train()
eng = train_engine.engine_for(net, B, 5)
obs = torch.randint(0, 2, (B, 9, 9, 9), device='cuda').float()
hid = torch.rand((B, 128, 9, 9), device='cuda')
act = torch.randint(0, 82, (B,), device='cuda')
g = torch.randn((B, 128, 9, 9), device='cuda')


def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


eng.begin_step()
layers = 2 * blocks
t_f = timed(lambda: eng.forward(2, 0, hid, None), reps)
# graph the same launch chain: what a captured training step sees (no host launch gaps)
def graphed(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            fn()
    return gr
gf = graphed(lambda: eng.forward(2, 0, hid, None))
t_fg = timed(gf.replay, reps)
def bwd():
    eng.backward(2, 0, g)
    eng.join()
gb = graphed(bwd)
t_bg = timed(gb.replay, reps)
print('prediction tower, %d conv layers, batch %d: forward %.1f us eager / %.1f us graph = %.2f us per conv + bn; '
      'backward %.1f us graph = %.2f us per layer (bn, dgrad; wgrad on its own stream)' % (layers, B, t_f, t_fg, t_fg / layers, t_bg, t_bg / layers))
