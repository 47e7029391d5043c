"""Layer-by-layer comparison of the training kernels (csrc/train.cu) with PyTorch autograd in fp32 on the same weights.
usage: python tools/train_check.py [blocks] [batch] [board]"""
import copy
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import muzero_b200 as mz
from muzero_b200 import train_engine
from muzero_b200.training import calc_loss_tensors, synthetic_transitions, _to_device

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
N = int(sys.argv[3]) if len(sys.argv) > 3 else 9
T = 5


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30)), float((a - b).abs().max()), float(b.abs().max())


torch.manual_seed(0)
net = mz.MuZeroBoardGameNet((9, N, N), N * N + 1, blocks, 128).cuda().train()
# make BatchNorm parameters non-trivial
with torch.no_grad():
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.uniform_(0.5, 1.5)
            m.bias.uniform_(-0.3, 0.3)
ref = copy.deepcopy(net)
tr, w = synthetic_transitions(net, B, T, seed=3)
inputs = _to_device(tr, 'cuda') + (torch.as_tensor(w).cuda(),)

# ---- reference: autograd modules, fp32 ----
os.environ['MZ_TRAIN_NATIVE'] = '0'
acts = {}


def hook(name):
    def f(mod, inp, out):
        acts.setdefault(name, []).append(out.detach())
    return f


for name, m in ref.named_modules():
    if isinstance(m, (torch.nn.Conv2d, mz.network.ResNetBlock)) or name.endswith('conv_block'):
        m.register_forward_hook(hook(name))
loss_r, pri_r = calc_loss_tensors(ref, *inputs)
loss_r.backward()

# ---- engine ----
os.environ['MZ_TRAIN_NATIVE'] = '1'
loss_e, pri_e = calc_loss_tensors(net, *inputs)
eng = train_engine.engine_for(net, B, T)
torch.cuda.synchronize()
print('loss ref %.6f engine %.6f  rel %.2e' % (loss_r.item(), loss_e.item(), abs(loss_e.item() - loss_r.item()) / abs(loss_r.item())))
towers = [('represent_net', 0, True), ('dynamics_net', 1, True), ('prediction_net', 2, False)]
for tname, tid, has_first in towers:
    for call in ([0] if tid == 0 else [0, T - 1]):
        layer = 0
        if has_first:
            y = eng.debug_view(tid, call, 0, 1)
            print(tname, call, 'conv_block.0 raw ', 'relL2 %.2e maxerr %.2e scale %.2e' % rel(y, acts[f'{tname}.conv_block.0'][call]))
            a = eng.debug_view(tid, call, 0, 2)
            print(tname, call, 'conv_block   act ', 'relL2 %.2e maxerr %.2e scale %.2e' % rel(a, acts[f'{tname}.conv_block'][call]))
            layer = 1
        for b in range(blocks):
            y1 = eng.debug_view(tid, call, layer, 1)
            print(tname, call, f'block{b} conv1 raw ', 'relL2 %.2e maxerr %.2e scale %.2e' % rel(y1, acts[f'{tname}.res_blocks.{b}.conv_block1.0'][call]))
            y2 = eng.debug_view(tid, call, layer + 1, 1)
            print(tname, call, f'block{b} conv2 raw ', 'relL2 %.2e maxerr %.2e scale %.2e' % rel(y2, acts[f'{tname}.res_blocks.{b}.conv_block2.0'][call]))
            a2 = eng.debug_view(tid, call, layer + 1, 2)
            print(tname, call, f'block{b} out       ', 'relL2 %.2e maxerr %.2e scale %.2e' % rel(a2, acts[f'{tname}.res_blocks.{b}'][call]))
            layer += 2
loss_e.backward()
torch.cuda.synchronize()
worst = 0.0
for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
    r = rel(p.grad, q.grad)
    worst = max(worst, r[0])
    print('grad %-55s relL2 %.2e maxerr %.2e scale %.2e' % (k, *r))
for (k, p), (_, q) in zip(net.named_buffers(), ref.named_buffers()):
    if p.dtype.is_floating_point:
        r = rel(p, q)
        if r[0] > 1e-3:
            print('buffer %-53s relL2 %.2e maxerr %.2e scale %.2e' % (k, *r))
    elif not torch.equal(p, q):
        print('buffer', k, p.item(), q.item())
print('worst grad relL2 %.3e' % worst)
