import copy, os, sys, torch
sys.path.insert(0, '/root/repo')
import muzero_b200 as mz
from muzero_b200.training import calc_loss, synthetic_transitions
os.environ['MZ_TRAIN_NATIVE'] = '0'
for board, cin, batch, unroll, blocks in [(9, 9, 128, 5, 8), (9, 3, 7, 5, 2), (9, 9, 128, 5, 2)]:
    torch.manual_seed(board * 100 + batch)
    A = board * board + 1
    net = mz.MuZeroBoardGameNet((cin, board, board), A, blocks, 128).cuda().train()
    twin = copy.deepcopy(net)
    tr, w = synthetic_transitions(net, batch, unroll, seed=batch)
    wt = torch.from_numpy(w).cuda()
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
    loss, _ = calc_loss(net, 'cuda', tr, wt); loss.backward()
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    ref, _ = calc_loss(twin, 'cuda', tr, wt); ref.backward()
    ratios = [float(p.grad.norm()) / (float(q.grad.norm()) + 1e-30) for p, q in zip(net.parameters(), twin.parameters())]
    print('TF32 autograd vs fp32 autograd: board %d batch %d blocks %d: loss rel %.1e, grad-norm ratio %.3f .. %.3f' % (board, batch, blocks, abs(float(loss)-float(ref))/abs(float(ref)), min(ratios), max(ratios)))
