"""%globaltimer timeline of one tower's forward and backward launch chain (MZ_TRAIN_TIMELINE=1): for every chain kernel
the gap since its predecessor finished, the time it sat in griddepcontrol.wait and its body time.
usage: MZ_TRAIN_TIMELINE=1 python tools/train_timeline.py [blocks] [batch]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

os.environ.setdefault('MZ_TRAIN_TIMELINE', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import muzero_b200 as mz
from muzero_b200 import _lib, train_engine

blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 8
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
torch.manual_seed(0)
net = mz.MuZeroBoardGameNet((9, 9, 9), 82, blocks, 128).cuda().train()
eng = train_engine.engine_for(net, B, 5)
hid = torch.rand((B, 128, 9, 9), device='cuda')
g = torch.randn((B, 128, 9, 9), device='cuda')
cudart = C.CDLL('libcudart.so.12')


def records():
    p, n, pr, fr = C.c_void_p(), C.c_size_t(), C.c_int32(), C.c_int32()
    _lib.check(_lib.lib().mz_train_debug_view(eng.handle, 2, 0, 0, 4, C.byref(p), C.byref(n), C.byref(pr), C.byref(fr)))
    host = np.zeros(n.value // 8, np.uint64)
    torch.cuda.synchronize()
    assert cudart.cudaMemcpy(C.c_void_p(host.ctypes.data), p, C.c_size_t(n.value), 2) == 0
    return host.reshape(-1, 10).astype(np.int64)


def graphed(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            fn()
    return gr


def show(title, rec, names):
    print(title)
    print('%-4s %-10s %9s %9s %9s %9s | cta 0: %9s %9s %9s %9s %9s' % ('#', 'kernel', 'gap us', 'wait us', 'body us', 'cta0 us', 'A tile', 'MMA issue', 'MMA drain', 'epilogue', 'sums+exit'))
    prev_end = None
    tot = {}
    for i, (t0, t1, t2, t3, t4, t5, t6, t7, kind, _) in enumerate(rec):
        gap = (t1 - prev_end) / 1e3 if prev_end is not None else float('nan')
        name = {1: 'conv', 2: 'dgrad', 3: 'bn_fwd', 4: 'bn_bwd', 5: 'wgrad'}.get(int(kind), '?')
        if name == 'wgrad':
            print('%-4d %-10s start %9.2f  end %9.2f  (%.2f us) | cta 0: first stage %.2f, MMA issue %.2f, drain %.2f, epilogue %.2f, exit %.2f' % (i, name, (t0 - rec[0][0]) / 1e3, (t2 - rec[0][0]) / 1e3, (t2 - t0) / 1e3, (t4 - t0) / 1e3, (t5 - t4) / 1e3, (t6 - t5) / 1e3, (t7 - t6) / 1e3, (t3 - t7) / 1e3))
            a = tot.setdefault(name, [0, 0.0, 0.0]); a[0] += 1; a[2] += (t2 - t0) / 1e3
            continue
        extra = ''
        if t4 and t7:
            extra = ' |        %9.2f %9.2f %9.2f %9.2f %9.2f' % ((t4 - t1) / 1e3, (t5 - t4) / 1e3, (t6 - t5) / 1e3, (t7 - t6) / 1e3, (t3 - t7) / 1e3)
        print('%-4d %-10s %9.2f %9.2f %9.2f %9.2f%s   [%.2f .. %.2f]' % (i, name, gap, (t1 - t0) / 1e3, (t2 - t1) / 1e3, (t3 - t1) / 1e3, extra, (t1 - rec[0][0]) / 1e3, (t2 - rec[0][0]) / 1e3))
        if prev_end is not None:
            a = tot.setdefault(name, [0, 0.0, 0.0])
            a[0] += 1; a[1] += gap; a[2] += (t2 - t1) / 1e3
        prev_end = t2
    for k, (n, gp, bd) in tot.items():
        print('  %-10s n=%d mean gap %.2f us, mean body %.2f us' % (k, n, gp / n, bd / n))
    print('  span %.1f us' % ((rec[-1][2] - rec[0][0]) / 1e3))


def fwd():
    eng.begin_step()
    eng.forward(2, 0, hid, None)


gf = graphed(fwd)
for _ in range(3):
    gf.replay()
show('forward (prediction tower): conv, bn alternate', records(), ['conv', 'bn'])


def bwd():
    eng.backward(2, 0, g)
    eng.join()


fwd()
# the backward's records follow the forward's in the same step
nf = len(records())
gb = graphed(bwd)
for _ in range(3):
    gb.replay()
rec = records()
show('backward (prediction tower): records after the forward chain', rec[nf:], [])
