"""Does a tree kernel of one sub-batch really run WHILE the persistent conv tower of the other one holds every SM?
Times (CUDA events) the tower alone, the fused expand+backup+select launch alone, and both launched together on two
streams in either order.  usage: python tools/overlap_probe.py [trees_per_part]"""
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

per = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
spec = bench.workload_spec('gomoku', per)
cfg = spec['cfg']
net = mz.MuZeroBoardGameNet(**spec['net_kw'])
net.load_state_dict(bench.state_dict_for(spec))
net = net.cuda().eval()
lib = _lib.lib()
plans = [SearchPlan(net, cfg, per, instance=i) for i in range(2)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
obs, mask, cur, opp = bench.synthetic_inputs(spec, per, 99)
for p in plans:
    p.use_graph = False
    p.pool.seed(1234 + np.arange(per))
    mz.uct_search_batch(obs, net, cfg, 1.0, mask, cur, opp, plan=p)      # leaves a fully grown tree behind
torch.cuda.synchronize()


def ptrs(p):
    pool = p.pool
    return dict(h=pool.hidden.data_ptr(), src=pool.view('SRC_SLOT').data_ptr(), dst=pool.view('DST_SLOT').data_ptr(),
                act=pool.view('LEAF_ACTION').data_ptr(), rew=pool.view('REWARD').data_ptr(),
                val=pool.view('VALUE').data_ptr())


def restart(p, sims):
    """fresh search state with `sims` simulations done, a select pending (what the loop looks like mid-search)"""
    s = _lib.current_stream()
    eng = net.engine(per, p.instance)
    q = ptrs(p)
    p.obs.copy_(torch.as_tensor(obs, device='cuda').reshape(per, -1).float())
    _lib.check(lib.mz_net_initial(eng['handle'], per, p.obs.data_ptr(), q['h'], p.root_slots.data_ptr(),
                                  p.pi0.data_ptr(), p.v0.data_ptr(), s))
    _lib.check(lib.mz_search_reset(p.pool.handle, p.pi0.data_ptr(), None, 0.0, None, p.players.data_ptr(), None, s))
    _lib.check(lib.mz_select(p.pool.handle, s))
    for _ in range(sims):
        _lib.check(lib.mz_net_recurrent(eng['handle'], per, q['h'], q['src'], q['act'], q['h'], q['dst'], q['rew'],
                                        q['val'], None, s))
        _lib.check(lib.mz_expand_backup_select(p.pool.handle, None, None, s))
    torch.cuda.synchronize()


def tower(p, s):
    eng = net.engine(per, p.instance)
    q = ptrs(p)
    _lib.check(lib.mz_net_recurrent(eng['handle'], per, q['h'], q['src'], q['act'], q['h'], q['dst'], q['rew'],
                                    q['val'], None, s.cuda_stream))


def tree(p, s):
    _lib.check(lib.mz_expand_backup_select(p.pool.handle, None, None, s.cuda_stream))


def ev():
    return torch.cuda.Event(enable_timing=True)


SIMS = 120
for trial in range(2):
    restart(plans[0], SIMS)
    restart(plans[1], SIMS)
    with torch.cuda.stream(streams[1]):
        tower(plans[1], streams[1])          # part 1 now has a tower result waiting for its tree kernel
    torch.cuda.synchronize()
    # --- alone
    a0, a1 = ev(), ev()
    a0.record(streams[0]); tower(plans[0], streams[0]); a1.record(streams[0])
    torch.cuda.synchronize()
    t_tower = a0.elapsed_time(a1) * 1e3
    b0, b1 = ev(), ev()
    b0.record(streams[1]); tree(plans[1], streams[1]); b1.record(streams[1])
    torch.cuda.synchronize()
    t_tree = b0.elapsed_time(b1) * 1e3
    # --- together, tower first
    with torch.cuda.stream(streams[0]):
        tree(plans[0], streams[0])
    with torch.cuda.stream(streams[1]):
        tower(plans[1], streams[1])
    torch.cuda.synchronize()
    g0, a1, b1 = ev(), ev(), ev()
    g0.record(streams[0]); streams[1].wait_event(g0)
    tower(plans[0], streams[0]); a1.record(streams[0])
    tree(plans[1], streams[1]); b1.record(streams[1])
    torch.cuda.synchronize()
    print(f'trial {trial}: tower alone {t_tower:.0f} us, tree kernel alone {t_tree:.0f} us | tower first: tower done '
          f'{g0.elapsed_time(a1) * 1e3:.0f} us, tree kernel done {g0.elapsed_time(b1) * 1e3:.0f} us')
    # --- together, tree kernel first
    with torch.cuda.stream(streams[0]):
        tree(plans[0], streams[0])
    with torch.cuda.stream(streams[1]):
        tower(plans[1], streams[1])
    torch.cuda.synchronize()
    g0, a1, b1 = ev(), ev(), ev()
    g0.record(streams[1]); streams[0].wait_event(g0)
    tree(plans[1], streams[1]); b1.record(streams[1])
    tower(plans[0], streams[0]); a1.record(streams[0])
    torch.cuda.synchronize()
    tf = (g0.elapsed_time(a1) * 1e3, g0.elapsed_time(b1) * 1e3)
    # --- a kernel with no shared memory and few registers beside the tower
    x = torch.zeros(1 << 20, device='cuda')
    torch.cuda.synchronize()
    g0, a1, b1 = ev(), ev(), ev()
    g0.record(streams[0]); streams[1].wait_event(g0)
    tower(plans[0], streams[0]); a1.record(streams[0])
    with torch.cuda.stream(streams[1]):
        x.add_(1.0)
    b1.record(streams[1])
    torch.cuda.synchronize()
    print(f'          tower first, then a torch elementwise kernel: tower done {g0.elapsed_time(a1) * 1e3:.0f} us, '
          f'elementwise done {g0.elapsed_time(b1) * 1e3:.0f} us')
    with torch.cuda.stream(streams[0]):
        tree(plans[0], streams[0])
    torch.cuda.synchronize()
    print(f'          tree kernel first: tower done {tf[0]:.0f} us, tree kernel done {tf[1]:.0f} us')
