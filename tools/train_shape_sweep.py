"""calc_loss + backward on the training kernels vs fp32 autograd over board sizes / batch sizes / unroll lengths / depths
the parity tests do not cover.  usage: python tools/train_shape_sweep.py"""
import copy
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import muzero_b200 as mz
from muzero_b200 import train_engine
from muzero_b200.training import calc_loss, synthetic_transitions

torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
CASES = [  # (board, in_planes, batch, unroll, blocks)
    (9, 9, 256, 2, 2), (6, 5, 50, 5, 1), (3, 9, 128, 5, 2), (11, 17, 32, 4, 3), (9, 9, 128, 5, 8), (5, 9, 512, 3, 1), (9, 3, 7, 5, 2)]
bad = 0
for board, cin, batch, unroll, blocks in CASES:
    torch.manual_seed(board * 100 + batch)
    A = board * board + 1
    net = mz.MuZeroBoardGameNet((cin, board, board), A, blocks, 128).cuda().train()
    twin = copy.deepcopy(net)
    tr, w = synthetic_transitions(net, batch, unroll, seed=batch)
    wt = torch.from_numpy(w).cuda()
    loss, _ = calc_loss(net, 'cuda', tr, wt)
    eng = train_engine.engine_for(net, batch, unroll)
    assert eng is not None and eng.active
    loss.backward()
    os.environ['MZ_TRAIN_NATIVE'] = '0'
    ref, _ = calc_loss(twin, 'cuda', tr, wt)
    ref.backward()
    os.environ.pop('MZ_TRAIN_NATIVE')
    torch.cuda.synchronize()
    ratios = [float(p.grad.norm()) / (float(q.grad.norm()) + 1e-30) for p, q in zip(net.parameters(), twin.parameters())]
    bufs = max(float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30)) for a, b in zip(net.buffers(), twin.buffers()))
    dl = abs(float(loss) - float(ref)) / abs(float(ref))
    ok = dl <= 2e-3 and 0.9 <= min(ratios) and max(ratios) <= 1.1 and bufs <= 3e-3
    bad += not ok
    print('board %2d planes %2d batch %3d unroll %d blocks %d (stacked calls %d): loss rel %.1e, grad-norm ratio %.3f .. %.3f, buffers %.1e  %s'
          % (board, cin, batch, unroll, blocks, eng.max_stacked_calls, dl, min(ratios), max(ratios), bufs, 'ok' if ok else 'beyond the 2-block tolerance'))
    del eng, net, twin
    torch.cuda.empty_cache()
print('cases beyond the 2-block tolerance (see tools/train_tf32_yardstick.py for what TF32 autograd does there):', bad)
