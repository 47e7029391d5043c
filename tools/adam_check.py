import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '/root/repo')
from muzero_b200 import _lib
torch.manual_seed(0)
lib = _lib.lib()
ps = [torch.randn(n, device='cuda') for n in (10000, 4096, 5, 300000)]
gs = [torch.randn_like(p) * 1e-3 for p in ps]
ms = [torch.randn_like(p) * 1e-4 for p in ps]
vs = [torch.rand_like(p) * 1e-7 for p in ps]
step = torch.tensor(3.0, device='cuda'); lr = torch.tensor(1e-3, device='cuda')
b1, b2, eps, wd = 0.9, 0.999, 1e-8, 1e-4
ref = []
for p, g, m, v in zip(ps, gs, ms, vs):
    p64, g64, m64, v64 = p.double(), g.double(), m.double(), v.double()
    g64 = g64 + wd * p64
    m64 = m64 + (1 - b1) * (g64 - m64)
    v64 = b2 * v64 + (1 - b2) * g64 * g64
    bc1, bc2 = 1 - b1 ** 3, 1 - b2 ** 3
    ref.append((p64 - (1e-3 / bc1) * m64 / (v64.sqrt() / bc2 ** 0.5 + eps), m64, v64))
chunk = lib.mz_adam_chunk_elements()
table = np.zeros((len(ps), 5), np.int64); ct, cs = [], []
for k, (p, g, m, v) in enumerate(zip(ps, gs, ms, vs)):
    table[k] = (p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel())
    for s in range(0, p.numel(), chunk):
        ct.append(k); cs.append(s)
tt = torch.from_numpy(table).cuda(); ctt = torch.tensor(ct, dtype=torch.int32, device='cuda'); cst = torch.tensor(cs, dtype=torch.int64, device='cuda')
_lib.check(lib.mz_adam_step(tt.data_ptr(), ctt.data_ptr(), cst.data_ptr(), len(ct), step.data_ptr(), lr.data_ptr(), b1, b2, eps, wd, _lib.current_stream()))
torch.cuda.synchronize()
for k, (p, m, v) in enumerate(zip(ps, ms, vs)):
    rp, rm, rv = ref[k]
    print(k, p.numel(), 'p', float((p.double() - rp).abs().max()), 'm', float((m.double() - rm).abs().max() / rm.abs().max()), 'v', float((v.double() - rv).abs().max() / rv.abs().max()))

# ---- against torch.optim.Adam(fused=True, capturable=True) itself, second step
import copy
pa = [torch.nn.Parameter(torch.randn(n, device='cuda')) for n in (10000, 777, 300000)]
pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
def make(params):
    return torch.optim.Adam(params, lr=torch.tensor(1e-3, device='cuda'), weight_decay=1e-4, capturable=True, fused=True)
oa, ob = make(pa), make(pb)
for it in range(3):
    gs = [torch.randn_like(p) * 1e-3 for p in pa]
    for p, q, g in zip(pa, pb, gs):
        p.grad = g.clone(); q.grad = g.clone()
    ob.step()
    if it == 0:
        oa.step()
    else:
        g0 = oa.param_groups[0]
        sts = [oa.state[p] for p in pa]
        torch._foreach_add_([s['step'] for s in sts], 1)
        table = np.zeros((len(pa), 5), np.int64); ct, cs = [], []
        for k, (p, s) in enumerate(zip(pa, sts)):
            table[k] = (p.data_ptr(), p.grad.data_ptr(), s['exp_avg'].data_ptr(), s['exp_avg_sq'].data_ptr(), p.numel())
            for st in range(0, p.numel(), chunk):
                ct.append(k); cs.append(st)
        tt = torch.from_numpy(table).cuda(); ctt = torch.tensor(ct, dtype=torch.int32, device='cuda'); cst = torch.tensor(cs, dtype=torch.int64, device='cuda')
        _lib.check(lib.mz_adam_step(tt.data_ptr(), ctt.data_ptr(), cst.data_ptr(), len(ct), sts[0]['step'].data_ptr(), g0['lr'].data_ptr(),
                                    g0['betas'][0], g0['betas'][1], g0['eps'], g0['weight_decay'], _lib.current_stream()))
    torch.cuda.synchronize()
    print('step', it + 1, 'max |p_a - p_b|', max(float((p - q).abs().max()) for p, q in zip(pa, pb)),
          'steps', float(oa.state[pa[0]]['step']), float(ob.state[pb[0]]['step']), 'lr type', type(oa.param_groups[0]['lr']))
