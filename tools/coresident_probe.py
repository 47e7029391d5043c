"""Footprint sweep: when do the blocks of a small kernel start if it is launched right after the persistent conv
tower (other stream)?  Needs tools/bin/libcoresident_probe.so (see tools/coresident_probe.cu)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import muzero_b200 as mz  # noqa: E402
from muzero_b200 import _lib  # noqa: E402
from muzero_b200.mcts import SearchPlan  # noqa: E402

pl = C.CDLL(os.path.join(ROOT, 'tools', 'bin', 'libcoresident_probe.so'))
pl.probe_launch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
pl.stamp_launch.argtypes = [C.c_void_p, C.c_void_p]
per = 1024
spec = bench.workload_spec('gomoku', per)
cfg = spec['cfg']
net = mz.MuZeroBoardGameNet(**spec['net_kw'])
net.load_state_dict(bench.state_dict_for(spec))
net = net.cuda().eval()
lib = _lib.lib()
plan = SearchPlan(net, cfg, per)
plan.use_graph = False
plan.pool.seed(1234 + np.arange(per))
obs, mask, cur, opp = bench.synthetic_inputs(spec, per, 99)
mz.uct_search_batch(obs, net, cfg, 1.0, mask, cur, opp, plan=plan)
torch.cuda.synchronize()
pool = plan.pool
eng = net.engine(per, 0)
q = dict(h=pool.hidden.data_ptr(), src=pool.view('SRC_SLOT').data_ptr(), dst=pool.view('DST_SLOT').data_ptr(),
         act=pool.view('LEAF_ACTION').data_ptr(), rew=pool.view('REWARD').data_ptr(), val=pool.view('VALUE').data_ptr())
s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()
BLOCKS = 256
out = torch.zeros(BLOCKS * 4, dtype=torch.int64, device='cuda')
st = torch.zeros(2, dtype=torch.int64, device='cuda')


def tower():
    _lib.check(lib.mz_net_recurrent(eng['handle'], per, q['h'], q['src'], q['act'], q['h'], q['dst'], q['rew'],
                                    q['val'], None, s0.cuda_stream))


for _ in range(3):
    tower()
torch.cuda.synchronize()
for regs_class, regs in ((0, 14), (1, 60), (2, 96)):
    for smem in (0, 1616, 3000):
        for rep in range(2):
            out.zero_()
            torch.cuda.synchronize()
            g0 = torch.cuda.Event()
            pl.stamp_launch(st.data_ptr(), s0.cuda_stream)
            g0.record(s0)
            s1.wait_event(g0)
            tower()
            pl.stamp_launch(st.data_ptr() + 8, s0.cuda_stream)
            pl.probe_launch(regs_class, BLOCKS, smem, 2000, out.data_ptr(), s1.cuda_stream)
            torch.cuda.synchronize()
        o = out.cpu().numpy().reshape(BLOCKS, 4)
        t = st.cpu().numpy()
        print(f'regs {regs:3d} smem {smem:5d}: tower {0} .. {(t[1] - t[0]) / 1e3:7.1f} us | probe blocks start '
              f'{(o[:, 0].min() - t[0]) / 1e3:7.1f} .. {(o[:, 0].max() - t[0]) / 1e3:7.1f} us, last end '
              f'{(o[:, 1].max() - t[0]) / 1e3:7.1f} us, on {len(set(o[:, 2]))} SMs, block length '
              f'{np.median(o[:, 1] - o[:, 0]) / 1e3:.1f} us')

# ---- dependent global loads beside the tower
pl.chase_launch.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
nxt = torch.randperm(1 << 20, device='cuda').int()
for with_tower in (False, True, True):
    out.zero_()
    torch.cuda.synchronize()
    g0 = torch.cuda.Event()
    pl.stamp_launch(st.data_ptr(), s0.cuda_stream)
    g0.record(s0)
    s1.wait_event(g0)
    if with_tower:
        tower()
    pl.stamp_launch(st.data_ptr() + 8, s0.cuda_stream)
    pl.chase_launch(BLOCKS, 64, out.data_ptr(), nxt.data_ptr(), s1.cuda_stream)
    torch.cuda.synchronize()
    o = out.cpu().numpy().reshape(BLOCKS, 4)
    t = st.cpu().numpy()
    print(f'chase 64 hops, tower={with_tower}: tower 0 .. {(t[1] - t[0]) / 1e3:.1f} us | blocks start '
          f'{(o[:, 0].min() - t[0]) / 1e3:.1f} .. {(o[:, 0].max() - t[0]) / 1e3:.1f} us, last end '
          f'{(o[:, 1].max() - t[0]) / 1e3:.1f} us, median block {np.median(o[:, 1] - o[:, 0]) / 1e3:.1f} us '
          f'= {np.median(o[:, 1] - o[:, 0]) / 64:.0f} ns per hop')

# ---- one instruction class at a time beside the tower
pl.opchain_launch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
tab = ((torch.arange(256, device='cuda') * 37 + 11) % 256).int()
big = torch.randint(0, 1 << 20, (200000 * 82 * 4,), device='cuda', dtype=torch.int32)
names = ['DFMA chain', 'REDUX+VOTE+SHFL chain', 'LDS chain', '__ldg chain (L1)', 'LDG.128 rows (262 MB)', 'DMUL/F2F/FADD/DADD']
iters = [4000, 2000, 4000, 2000, 100, 2000]
for mode in range(6):
    res = []
    for with_tower in (False, True):
        for rep in range(2):
            out.zero_()
            torch.cuda.synchronize()
            g0 = torch.cuda.Event()
            pl.stamp_launch(st.data_ptr(), s0.cuda_stream)
            g0.record(s0)
            s1.wait_event(g0)
            if with_tower:
                tower()
            pl.opchain_launch(mode, 148, iters[mode], out.data_ptr(), tab.data_ptr(), big.data_ptr(), s1.cuda_stream)
            torch.cuda.synchronize()
        o = out.cpu().numpy().reshape(BLOCKS, 4)[:148]
        res.append(np.median(o[:, 1] - o[:, 0]) / iters[mode])
    print(f'{names[mode]:28s}: {res[0]:8.1f} ns per step alone, {res[1]:8.1f} ns beside the tower ({res[1] / res[0]:.2f}x)')

# ---- the real tree kernel of a second sub-batch beside the tower
plan2 = SearchPlan(net, cfg, per, instance=1)
plan2.use_graph = False
plan2.pool.seed(99 + np.arange(per))
cfg2_sims = cfg.num_simulations
mz.uct_search_batch(obs, net, cfg, 1.0, mask, cur, opp, plan=plan2)
torch.cuda.synchronize()
# rebuild a mid-search state in plan2: reset + 100 simulations
p2 = plan2.pool
eng2 = net.engine(per, 1)
q2 = dict(h=p2.hidden.data_ptr(), src=p2.view('SRC_SLOT').data_ptr(), dst=p2.view('DST_SLOT').data_ptr(),
          act=p2.view('LEAF_ACTION').data_ptr(), rew=p2.view('REWARD').data_ptr(), val=p2.view('VALUE').data_ptr())
with torch.cuda.stream(s1):
    _lib.check(lib.mz_search_reset(p2.handle, plan2.pi0.data_ptr(), None, 0.0, None, plan2.players.data_ptr(), None,
                                   s1.cuda_stream))
    _lib.check(lib.mz_select(p2.handle, s1.cuda_stream))
    for _ in range(100):
        _lib.check(lib.mz_net_recurrent(eng2['handle'], per, q2['h'], q2['src'], q2['act'], q2['h'], q2['dst'],
                                        q2['rew'], q2['val'], None, s1.cuda_stream))
        _lib.check(lib.mz_expand_backup_select(p2.handle, None, None, s1.cuda_stream))
torch.cuda.synchronize()
st2 = torch.zeros(2, dtype=torch.int64, device='cuda')
for rep in range(3):
    with torch.cuda.stream(s1):          # a tower result for the tree kernel to consume
        _lib.check(lib.mz_net_recurrent(eng2['handle'], per, q2['h'], q2['src'], q2['act'], q2['h'], q2['dst'],
                                        q2['rew'], q2['val'], None, s1.cuda_stream))
    torch.cuda.synchronize()
    sv = p2.view('STATS')
    sv[4] = (1 << 62); sv[5] = 0; sv[6] = 0
    torch.cuda.synchronize()
    g0 = torch.cuda.Event()
    pl.stamp_launch(st.data_ptr(), s0.cuda_stream)
    g0.record(s0)
    s1.wait_event(g0)
    tower()
    pl.stamp_launch(st.data_ptr() + 8, s0.cuda_stream)
    pl.stamp_launch(st2.data_ptr(), s1.cuda_stream)
    _lib.check(lib.mz_expand_backup_select(p2.handle, None, None, s1.cuda_stream))
    pl.stamp_launch(st2.data_ptr() + 8, s1.cuda_stream)
    torch.cuda.synchronize()
    t, t2 = st.cpu().numpy(), st2.cpu().numpy()
    print(f'real tree kernel: tower 0 .. {(t[1] - t[0]) / 1e3:.1f} us | stamp before tree kernel at '
          f'{(t2[0] - t[0]) / 1e3:.1f} us, after it at {(t2[1] - t[0]) / 1e3:.1f} us')
    k = sv.cpu().numpy()
    if k[5]:
        print(f'   in-kernel (MZ_TREE_TIMING): first block starts {(k[4] - t[0]) / 1e3:.1f} us, last block starts '
              f'{(k[5] - t[0]) / 1e3:.1f} us, last tree done {(k[6] - t[0]) / 1e3:.1f} us')
