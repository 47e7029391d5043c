"""Verbose form of tests/test_train_gpu.py::test_towers_forward_backward_vs_autograd. usage: [blocks] [batch] [board]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from test_train_gpu import tower_errors
a = [int(x) for x in sys.argv[1:]] + [2, 16, 9][len(sys.argv) - 1:]
mine = tower_errors(*a[:3])
tf32 = tower_errors(*a[:3], tf32_autograd=True)
print('%-60s %-12s %s' % ('relative L2 error vs fp32 autograd', 'kernels', 'TF32 autograd'))
for k, v in mine.items():
    print('%-60s %.3e    %.3e' % (k, v, tf32.get(k, float('nan'))))
