"""GPU parity tests of the hand-written training towers (csrc/train.cu, muzero_b200/train_engine.py): every tower
forward / backward against PyTorch autograd in fp32 on the same weights, inputs and output gradients, and the whole
K-step unroll (`calc_loss`) against the reference's recorded loss and gradients.

Stated tolerances (fp16 activations and weights, bf16 gradients, fp32 accumulation; fp32 autograd without TF32 as the
yardstick): relative L2 error per tensor, forward <= 3e-3, gradients <= 1e-1 each and <= 5e-2 in the median.  The gradient figure is not rounding
noise of the backward kernels: a forward error of 1e-3 flips the ReLU mask of about one activation in a thousand, and
each flip adds or removes that element's whole gradient -- sqrt(1e-3) = 3 % of a gradient's norm (measured 2-4 %; PyTorch's
own TF32 convolutions, which also carry 10 mantissa bits, sit at the same distance from fp32:
tools/train_tower_check.py prints that figure beside the kernels')."""
import copy
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

FWD_TOL, GRAD_TOL, GRAD_MEDIAN_TOL = 3e-3, 1e-1, 5e-2


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(((a - b).norm() / (b.norm() + 1e-30)).detach())


def tower_errors(blocks=2, batch=16, board=9, seed=0, in_planes=9, tf32_autograd=False):
    """Per-tower comparison: {name: relative L2 error} for outputs, input gradients and every parameter gradient."""
    import muzero_b200 as mz
    from muzero_b200 import train_engine
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        torch.manual_seed(seed)
        A = board * board + 1
        net = mz.MuZeroBoardGameNet((in_planes, board, board), A, blocks, 128).cuda().train()
        with torch.no_grad():
            for m in net.modules():
                if isinstance(m, torch.nn.BatchNorm2d):
                    m.weight.uniform_(0.5, 1.5)
                    m.bias.uniform_(-0.3, 0.3)
        ref = copy.deepcopy(net)
        gen = torch.Generator(device='cuda').manual_seed(seed + 1)
        obs = torch.randint(0, 2, (batch, in_planes, board, board), device='cuda', generator=gen).float()
        hid = torch.rand((batch, 128, board, board), device='cuda', generator=gen)
        act = torch.randint(0, A, (batch, 1), device='cuda', generator=gen)
        gouts = [torch.randn((batch, 128, board, board), device='cuda', generator=gen) * s for s in (1.0, 0.3, 2.0)]
        # reference: the autograd modules in fp32
        h1 = hid.clone().requires_grad_(True)
        h2 = hid.clone().requires_grad_(True)
        planes = mz.network.action_planes(act, A, board, board)
        r_out = [ref.represent_net.res_blocks(ref.represent_net.conv_block(obs)),
                 ref.dynamics_net.res_blocks(ref.dynamics_net.conv_block(torch.cat([h1, planes], dim=1))),
                 ref.prediction_net.res_blocks(h2)]
        torch.autograd.backward(r_out, gouts)
        e1 = hid.clone().requires_grad_(True)
        e2 = hid.clone().requires_grad_(True)
        if tf32_autograd:
            # yardstick: the same autograd modules with PyTorch's TF32 convolutions instead of the kernels
            torch.backends.cudnn.allow_tf32 = True
            e_out = [net.represent_net.res_blocks(net.represent_net.conv_block(obs)),
                     net.dynamics_net.res_blocks(net.dynamics_net.conv_block(torch.cat([e1, planes], dim=1))),
                     net.prediction_net.res_blocks(e2)]
            torch.autograd.backward(e_out, gouts)
            eng = type('E', (), {'modules': train_engine._tower_modules(net)})
        else:
            eng = train_engine.engine_for(net, batch, 5)
            assert eng is not None
            e_out = [train_engine.tower(eng, 0, obs), train_engine.tower(eng, 1, e1, act), train_engine.tower(eng, 2, e2)]
            for k in (2, 1, 0):                       # the representation tower's backward closes the step
                e_out[k].backward(gouts[k])
        torch.cuda.synchronize()
        errs = {}
        for k, name in enumerate(('represent', 'dynamics', 'prediction')):
            errs[f'out/{name}'] = rel(e_out[k], r_out[k])
        errs['grad_in/dynamics'] = rel(e1.grad, h1.grad)
        errs['grad_in/prediction'] = rel(e2.grad, h2.grad)
        tower_params = set()
        for conv, bn in eng.modules:
            tower_params.update([id(conv.weight), id(bn.weight), id(bn.bias)])
        for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
            if id(p) in tower_params:
                errs[f'grad/{k}'] = rel(p.grad, q.grad)
        for (k, p), (_, q) in zip(net.named_buffers(), ref.named_buffers()):
            if 'running' in k and not ('head' in k or 'policy' in k or 'value' in k):
                errs[f'buffer/{k}'] = rel(p, q)
        return errs
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


@pytest.mark.parametrize('blocks,batch,board,in_planes', [(2, 16, 9, 9), (1, 5, 5, 17), (3, 128, 9, 9)])
def test_towers_forward_backward_vs_autograd(blocks, batch, board, in_planes):
    errs = tower_errors(blocks, batch, board, in_planes=in_planes)
    bad = {k: v for k, v in errs.items()
           if v > (FWD_TOL if k.startswith(('out/', 'buffer/')) else GRAD_TOL) or not np.isfinite(v)}
    assert not bad, bad
    assert float(np.median([v for k, v in errs.items() if k.startswith('grad')])) <= GRAD_MEDIAN_TOL


def _golden_net():
    import muzero_b200 as mz
    torch.manual_seed(23)
    return mz.MuZeroBoardGameNet((9, 9, 9), 82, 2, 128)


def test_calc_loss_on_the_training_kernels_vs_reference_recording():
    """tests/golden/train_golden_r2.npz: what the unmodified reference's pipeline.calc_loss + backward produced for
    this network and batch (make_golden_train_r2.py).  Stated tolerance of the fp16 / bf16 kernels over the whole
    5-step unroll: loss 2e-3 relative, priorities 2e-2 absolute; per-parameter gradient norm within 10 % and direction
    (cosine over the recorded entries) >= 0.95 (measured 0.97-0.99 / within 4 %) -- gradients of an unrolled ReLU /
    min-max network are not Lipschitz in the activations (a flipped ReLU or arg-max re-routes them, see the module
    docstring), so entrywise bounds would only measure luck."""
    from muzero_b200 import train_engine
    from muzero_b200.training import calc_loss, synthetic_transitions
    z = np.load(os.path.join(GOLDEN, 'train_golden_r2.npz'))
    net = _golden_net().cuda().train()
    tr, w = synthetic_transitions(net, 16, 5, seed=78)
    loss, pri = calc_loss(net, 'cuda', tr, torch.from_numpy(w).cuda())
    assert train_engine.engine_for(net, 16, 5) is not None and train_engine.engine_for(net, 16, 5).active
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - float(z['loss'])) <= 2e-3 * abs(float(z['loss'])), (loss.item(), float(z['loss']))
    np.testing.assert_allclose(pri, z['priorities'], rtol=0, atol=2e-2)
    bad = {}
    for k, p in net.named_parameters():
        g = p.grad.detach().reshape(-1).double().cpu()
        want_norm, want_head = float(z[f'grad_norm_{k}']), torch.from_numpy(z[f'grad_head_{k}']).double()
        head = g[:want_head.numel()]
        cos = float((head * want_head).sum() / (head.norm() * want_head.norm() + 1e-30))
        nrm = float(g.norm()) / (want_norm + 1e-30)
        if not (0.9 <= nrm <= 1.1 and cos >= 0.95):
            bad[k] = (round(nrm, 4), round(cos, 4))
    assert not bad, bad
    for k, b in net.named_buffers():
        if 'running' in k:
            np.testing.assert_allclose(b.cpu().numpy(), z[f'buffer_{k}'], rtol=2e-2, atol=2e-3, err_msg=k)


def test_native_training_step_under_the_learner_graph():
    """DataParallelLearner with the training kernels: eager iterations and CUDA-graph replays agree step by step, the
    weights move, and the inference engine sees them."""
    import muzero_b200 as mz
    from muzero_b200.training import DataParallelLearner, synthetic_transitions
    torch.manual_seed(0)
    net_a = mz.MuZeroBoardGameNet((9, 9, 9), 82, 2, 128).cuda()
    net_b = copy.deepcopy(net_a)
    cfg = mz.config.make_gomoku_config(num_training_steps=10, batch_size=16)
    la = DataParallelLearner(net_a, cfg, 'cuda', use_graph=True)
    lb = DataParallelLearner(net_b, cfg, 'cuda', use_graph=False)
    assert la.native_towers and not la.channels_last
    before = net_a.represent_net.conv_block[0].weight.detach().clone()
    for it in range(6):
        tr, w = synthetic_transitions(net_a, 16, 5, seed=it)
        net_b.load_state_dict(net_a.state_dict())
        for sa, sb in zip(la.optimizer.state.values(), lb.optimizer.state.values()):
            for k in sa:
                sb[k].copy_(sa[k])
        loss_a, pa = la.step(tr, w)
        loss_b, pb = lb.step(tr, w)
        # same kernels, same inputs, and nothing in them depends on scheduling (per-tile partial sums are added in a fixed
        # order, no floating-point atomics): the replayed graph reproduces the eager iteration
        assert abs(loss_a - loss_b) <= 1e-6 * max(1.0, abs(loss_b)), (it, loss_a, loss_b)
        gmax = float(lb.flat_grad.abs().max())
        assert float((la.flat_grad - lb.flat_grad).abs().max()) <= 1e-5 * gmax, (it, float((la.flat_grad - lb.flat_grad).abs().max()), gmax)
    assert la._graph is not None and lb._graph is None
    assert float((net_a.represent_net.conv_block[0].weight - before).abs().max()) > 1e-5
    net_a.eval()
    obs = torch.randint(0, 2, (4, 9, 9, 9), device='cuda').float()
    _, pi, v = net_a.initial_inference_batch(obs)
    assert torch.isfinite(pi).all() and torch.isfinite(v).all()


def test_stacked_prediction_calls_equal_separate_calls():
    """mz_train_tower_forward_calls / _backward_calls (the K prediction calls of an unroll as one launch chain) against K
    separate calls on the same inputs: per-call BatchNorm statistics, running statistics and parameter gradients in call
    order.  Nothing in the kernels depends on scheduling (fixed-point statistics, fixed summation orders), so outputs,
    input gradients, parameter gradients and buffers are the same BITS."""
    import muzero_b200 as mz
    from muzero_b200 import train_engine
    torch.manual_seed(5)
    K, B = 4, 64                                   # 64 * 100 rows per call: a multiple of 256
    net = mz.MuZeroBoardGameNet((9, 9, 9), 82, 2, 128).cuda().train()
    twin = copy.deepcopy(net)
    gen = torch.Generator(device='cuda').manual_seed(6)
    obs = torch.randint(0, 2, (B, 9, 9, 9), device='cuda', generator=gen).float()
    hids = [torch.rand((B, 128, 9, 9), device='cuda', generator=gen) * (k + 1) for k in range(K)]
    gouts = [torch.randn((B, 128, 9, 9), device='cuda', generator=gen) * (0.5 + k) for k in range(K)]
    results = []
    for model, stacked in ((net, True), (twin, False)):
        eng = train_engine.engine_for(model, B, 5)
        assert eng.max_stacked_calls >= K
        xs = [h.clone().requires_grad_(True) for h in hids]
        rep = train_engine.tower(eng, 0, obs)                       # opens the step
        if stacked:
            out = train_engine.prediction_calls(eng, xs)
            assert eng.calls[2] == K
        else:
            out = torch.cat([train_engine.tower(eng, 2, x) for x in xs], dim=0)
        out.backward(torch.cat(gouts, dim=0))
        rep.backward(torch.zeros_like(rep))                          # closes the step (weight gradients are finalised)
        torch.cuda.synchronize()
        results.append((out.detach(), [x.grad for x in xs], model))
    (o1, g1, m1), (o2, g2, m2) = results
    assert torch.equal(o1, o2)
    for a, b in zip(g1, g2):
        assert torch.equal(a, b)
    for (k, p), (_, q) in zip(m1.named_parameters(), m2.named_parameters()):
        if k.startswith('prediction_net.res_blocks'):
            assert p.grad is not None and torch.equal(p.grad, q.grad), k
    for (k, a), (_, b) in zip(m1.named_buffers(), m2.named_buffers()):
        if k.startswith('prediction_net.res_blocks'):
            assert torch.equal(a, b), k


def test_fused_hidden_state_normalisation_equals_the_torch_one():
    """tower(mode=2): the dynamics tower's output AND its min-max normalisation (util.py:31-36) from one kernel pass,
    with the normalisation's backward folded into the pass that builds the tower's output gradient -- against
    ``normalize_hidden_state`` applied by PyTorch to the raw output (what the autograd path does).  Forward: the same
    float32 operations on the same values, so the same bits.  Backward: both paths round the float32 gradient to bf16
    once; they differ by float32 summation order only (<= 2e-3 relative L2 stated, ~1e-4 measured)."""
    import muzero_b200 as mz
    from muzero_b200 import train_engine
    from muzero_b200.network import normalize_hidden_state
    torch.manual_seed(9)
    B = 16
    net = mz.MuZeroBoardGameNet((9, 9, 9), 82, 2, 128).cuda().train()
    twin = copy.deepcopy(net)
    gen = torch.Generator(device='cuda').manual_seed(10)
    obs = torch.randint(0, 2, (B, 9, 9, 9), device='cuda', generator=gen).float()
    hid = torch.rand((B, 128, 9, 9), device='cuda', generator=gen)
    act = torch.randint(0, 82, (B, 1), device='cuda', generator=gen)
    g_raw = torch.randn((B, 128, 9, 9), device='cuda', generator=gen)
    g_norm = torch.randn((B, 128, 9, 9), device='cuda', generator=gen) * 3.0
    got = []
    for model, fused in ((net, True), (twin, False)):
        eng = train_engine.engine_for(model, B, 5)
        x = hid.clone().requires_grad_(True)
        rep = train_engine.tower(eng, 0, obs, mode=1 if fused else 0)
        rep_n = rep if fused else normalize_hidden_state(rep)
        if fused:
            raw, nrm = train_engine.tower(eng, 1, x, act, mode=2)
        else:
            raw = train_engine.tower(eng, 1, x, act)
            nrm = normalize_hidden_state(raw)
        torch.autograd.backward([raw, nrm], [g_raw, g_norm])
        rep_n.backward(g_norm * 0.1)
        torch.cuda.synchronize()
        got.append((raw.detach(), nrm.detach(), rep_n.detach(), x.grad, model))
    (r1, n1, p1, gx1, m1), (r2, n2, p2, gx2, m2) = got
    assert torch.equal(r1, r2) and torch.equal(n1, n2) and torch.equal(p1, p2)
    assert float(n1.min()) >= 0.0 and float(n1.max()) <= 1.0
    assert rel(gx1, gx2) <= 2e-3, rel(gx1, gx2)
    worst = max(rel(p.grad, q.grad) for (k, p), (_, q) in zip(m1.named_parameters(), m2.named_parameters())
                if p.grad is not None and (k.startswith('dynamics_net.conv_block') or k.startswith('dynamics_net.res_blocks')
                                           or k.startswith('represent_net')))
    assert worst <= 5e-3, worst


def _native_vs_autograd(batch, unroll, blocks=2, seed=11, accumulate=1):
    """calc_loss + backward on the training kernels and, on a copy of the network, through fp32 autograd (no TF32)."""
    import muzero_b200 as mz
    from muzero_b200 import train_engine
    from muzero_b200.training import calc_loss, synthetic_transitions
    torch.manual_seed(seed)
    net = mz.MuZeroBoardGameNet((9, 9, 9), 82, blocks, 128).cuda().train()
    twin = copy.deepcopy(net)
    losses = []
    for k in range(accumulate):
        tr, w = synthetic_transitions(net, batch, unroll, seed=seed + 1 + k)
        wt = torch.from_numpy(w).cuda()
        loss, _ = calc_loss(net, 'cuda', tr, wt)
        eng = train_engine.engine_for(net, batch, unroll)
        assert eng is not None and eng.active
        loss.backward()
        prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
        os.environ['MZ_TRAIN_NATIVE'] = '0'
        try:
            ref, _ = calc_loss(twin, 'cuda', tr, wt)
            ref.backward()
        finally:
            os.environ.pop('MZ_TRAIN_NATIVE', None)
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
        losses.append((float(loss.detach()), float(ref.detach())))
    torch.cuda.synchronize()
    return net, twin, losses, eng


@pytest.mark.parametrize('batch,unroll', [(64, 3), (32, 6), (24, 5)])
def test_calc_loss_native_vs_autograd_other_shapes(batch, unroll):
    """Unroll lengths other than 5 and batch sizes with (64, 32: rows per call a multiple of 128) and without (24) stacked
    prediction calls: loss within 2e-3 of fp32 autograd, every gradient's norm within 10 %, BatchNorm buffers within 3e-3."""
    net, twin, losses, eng = _native_vs_autograd(batch, unroll)
    assert (eng.max_stacked_calls >= unroll) == (batch * 100 % 128 == 0)
    for got, want in losses:
        assert abs(got - want) <= 2e-3 * abs(want), (got, want)
    for (k, p), (_, q) in zip(net.named_parameters(), twin.named_parameters()):
        r = float(p.grad.norm()) / (float(q.grad.norm()) + 1e-30)
        assert 0.9 <= r <= 1.1, (k, r)
    for (k, a), (_, b) in zip(net.named_buffers(), twin.named_buffers()):
        if a.dtype.is_floating_point:
            assert rel(a, b) <= FWD_TOL, (k, rel(a, b))
        else:
            assert torch.equal(a, b), k


def test_gradients_accumulate_over_two_backward_passes():
    """Two calc_loss + backward passes without zeroing in between: the kernels add into .grad like autograd does
    (conv weights through wgrad_finalize_kernel, BatchNorm weights / biases in bn_bwd_apply_kernel)."""
    net, twin, losses, _ = _native_vs_autograd(32, 5, accumulate=2)
    single, _, _, _ = _native_vs_autograd(32, 5, accumulate=1)
    grew = 0
    for (k, p), (_, q), (_, s) in zip(net.named_parameters(), twin.named_parameters(), single.named_parameters()):
        r = float(p.grad.norm()) / (float(q.grad.norm()) + 1e-30)
        assert 0.9 <= r <= 1.1, (k, r)
        grew += float(p.grad.norm()) > 1.2 * float(s.grad.norm())
    assert grew >= 0.8 * len(list(net.parameters()))          # the second pass really was added


def test_one_launch_adam_equals_torch_adam(monkeypatch):
    """DataParallelLearner's optimizer step through csrc/optim.cu (one launch over a chunk table, on the torch optimizer's
    own exp_avg / exp_avg_sq / step tensors) against torch.optim.Adam(fused=True, capturable=True) on a twin learner fed
    the SAME gradients (whole training steps cannot be compared this tightly: the PyTorch head kernels are not
    bit-reproducible, and Adam turns a 1e-8 difference into a different step wherever a gradient is tiny).  Parameters
    and optimizer state agree to float32 rounding after every step (parameters: to 5e-5 of a step's size, see below),
    weight decay on."""
    import muzero_b200 as mz
    from muzero_b200.training import DataParallelLearner, synthetic_transitions
    torch.manual_seed(13)
    net_a = mz.MuZeroBoardGameNet((9, 9, 9), 82, 2, 128).cuda()
    net_b = copy.deepcopy(net_a)
    cfg = mz.config.make_gomoku_config(num_training_steps=100, batch_size=32)
    cfg.weight_decay = 1e-4
    monkeypatch.setenv('MZ_FAST_ADAM', '1')
    la = DataParallelLearner(net_a, cfg, 'cuda', use_graph=False)
    monkeypatch.setenv('MZ_FAST_ADAM', '0')
    lb = DataParallelLearner(net_b, cfg, 'cuda', use_graph=False)
    tr, w = synthetic_transitions(net_a, 32, 5, seed=40)
    monkeypatch.setenv('MZ_FAST_ADAM', '1')
    la.step(tr, w)                                   # creates the optimizer state (torch's step), real gradients
    assert la._fast_adam is None
    lb.flat_grad.copy_(la.flat_grad)
    with torch.no_grad():                            # the twin starts from the same parameters ...
        for p, q in zip(net_a.parameters(), net_b.parameters()):
            q.copy_(p)
    monkeypatch.setenv('MZ_FAST_ADAM', '0')
    lb.optimizer.step()                              # ... and gets a state of its own,
    for p, q in zip(la.params, lb.params):           # which is then made equal to the first learner's
        for key in ('exp_avg', 'exp_avg_sq', 'step'):
            lb.optimizer.state[q][key].copy_(la.optimizer.state[p][key])
        with torch.no_grad():
            q.copy_(p)
    gen = torch.Generator(device='cuda').manual_seed(3)
    for it in range(4):
        g = torch.randn(la.flat_grad.shape, device='cuda', generator=gen) * (10.0 ** -(2 * it))
        la.flat_grad.copy_(g)
        lb.flat_grad.copy_(g)
        monkeypatch.setenv('MZ_FAST_ADAM', '1')
        la._optimizer_step()
        assert isinstance(la._fast_adam, dict)
        lb.optimizer.step()
        torch.cuda.synchronize()
        for (k, p), (_, q) in zip(net_a.named_parameters(), net_b.named_parameters()):
            # torch forms the bias corrections 1 - beta^step in float32 (1 - 0.998001 loses 3e-5 of its value at step 2),
            # the kernel in double: the two steps differ by up to ~2e-5 of a step's size (lr = 2e-3)
            assert float((p - q).abs().max()) <= 2e-6 * float(q.abs().max()) + 1e-7, (it, k, float((p - q).abs().max()))
            sa, sb = la.optimizer.state[p], lb.optimizer.state[q]
            assert float(sa['step']) == float(sb['step']) == it + 2
            for key in ('exp_avg', 'exp_avg_sq'):
                d, m = float((sa[key] - sb[key]).abs().max()), float(sb[key].abs().max())
                assert d <= 5e-6 * (m + 1e-30), (it, k, key, d, m)


def test_head_conv_kernels_equal_conv2d():
    """mz_head_conv_forward / _backward (the heads' 1x1 convolutions over stacked tower outputs) against F.conv2d in fp32
    (no TF32): output, input gradient and weight gradient to float32 summation-order rounding, for the policy (2), value
    (1) and stacked policy + value (3) shapes."""
    from muzero_b200.network import _HeadConv1x1
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        gen = torch.Generator(device='cuda').manual_seed(21)
        for n, c, h, m in ((640, 128, 9, 3), (37, 128, 9, 1), (5, 64, 6, 2), (130, 256, 11, 4)):
            x = torch.randn((n, c, h, h), device='cuda', generator=gen)
            w = torch.randn((m, c, 1, 1), device='cuda', generator=gen) * 0.1
            dy = torch.randn((n, m, h, h), device='cuda', generator=gen)
            xa, wa = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
            xb, wb = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
            ya = _HeadConv1x1.apply(xa, wa)
            yb = torch.nn.functional.conv2d(xb, wb)
            ya.backward(dy)
            yb.backward(dy)
            torch.cuda.synchronize()
            assert rel(ya, yb) <= 2e-6 and rel(xa.grad, xb.grad) <= 2e-6 and rel(wa.grad, wb.grad) <= 2e-5, \
                (n, c, h, m, rel(ya, yb), rel(xa.grad, xb.grad), rel(wa.grad, wb.grad))
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


def test_reference_learner_loop_on_the_training_kernels():
    """pipeline.py:232-257 as the reference writes it -- optimizer.zero_grad() (set_to_none, the torch default: the
    engine finds new gradient tensors every step and re-binds), calc_loss, backward, torch.optim.Adam.step() -- on a
    network whose towers run on csrc/train.cu: every parameter gets a finite gradient each step and the loss of a fixed
    batch falls."""
    import muzero_b200 as mz
    from muzero_b200 import train_engine
    from muzero_b200.training import calc_loss, synthetic_transitions
    torch.manual_seed(0)
    net = mz.MuZeroBoardGameNet((9, 9, 9), 82, 2, 128).cuda().train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    tr, w = synthetic_transitions(net, 64, 5, seed=1)
    wt = torch.from_numpy(w).cuda()
    losses = []
    for _ in range(6):
        opt.zero_grad()
        loss, _pri = calc_loss(net, 'cuda', tr, wt)
        assert train_engine.engine_for(net, 64, 5).active
        loss.backward()
        assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in net.parameters())
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < 0.8 * losses[0], losses


@pytest.mark.parametrize('mid,out,calls,batch', [(2, 82, 5, 32), (1, 1, 3, 16), (1, 1, 1, 7)])
def test_head_tail_kernels_equal_the_torch_head(monkeypatch, mid, out, calls, batch):
    """mz_head_tail_forward / _backward (a head's train-mode BatchNorm with per-call statistics + ReLU + Flatten + Linear over
    stacked calls) against the PyTorch form of network.head_over_calls on a twin head: outputs, input gradient, all four
    parameter gradients and the BatchNorm buffers to float32 rounding (1e-4 relative L2; the two sum in different orders)."""
    from muzero_b200.network import _head, _head_after_conv
    torch.manual_seed(31)
    ha = _head(128, mid, 81, out).cuda().train()
    with torch.no_grad():
        ha[1].weight.uniform_(0.5, 1.5)
        ha[1].bias.uniform_(-0.3, 0.3)
    hb = copy.deepcopy(ha)
    gen = torch.Generator(device='cuda').manual_seed(32)
    y = torch.randn((calls * batch, mid, 9, 9), device='cuda', generator=gen) * 2.0 + 0.5
    dout = torch.randn((calls * batch, out), device='cuda', generator=gen)
    ya, yb = y.clone().requires_grad_(True), y.clone().requires_grad_(True)
    monkeypatch.setenv('MZ_HEAD_TAIL', '1')
    oa = _head_after_conv(ha, ya, calls)
    monkeypatch.setenv('MZ_HEAD_TAIL', '0')
    ob = _head_after_conv(hb, yb, calls)
    oa.backward(dout)
    ob.backward(dout)
    torch.cuda.synchronize()
    assert rel(oa, ob) <= 1e-5, rel(oa, ob)
    assert rel(ya.grad, yb.grad) <= 1e-4, rel(ya.grad, yb.grad)
    for (k, p), (_, q) in zip(ha.named_parameters(), hb.named_parameters()):
        if p.grad is None and q.grad is None:
            continue                                  # the 1x1 convolution's weight is not part of the tail
        assert rel(p.grad, q.grad) <= 1e-4, (k, rel(p.grad, q.grad))
    for (k, a), (_, b) in zip(ha.named_buffers(), hb.named_buffers()):
        if a.dtype.is_floating_point:
            assert rel(a, b) <= 1e-5, (k, rel(a, b))
        else:
            assert torch.equal(a, b), k
