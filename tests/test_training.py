"""CPU tests of the K-step-unroll training step: `calc_loss` against the reference's own recorded
loss / priorities / gradients, and the data-parallel step over a world_size-2 gloo group."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN
import muzero_b200 as mz
from muzero_b200.training import (DataParallelLearner, Transition, calc_loss, scalar_to_categorical_probabilities,
                                  signed_hyperbolic, signed_parabolic, synthetic_transitions, transform_to_2hot)

CASES = {
    'tictactoe_mlp': ('mlp', dict(input_shape=(9, 3, 3), num_actions=10, num_planes=256, value_support_size=1,
                                  reward_support_size=1, hidden_dim=64), 'ckpt_tictactoe.npz', 16),
    'cartpole_mlp': ('mlp', dict(input_shape=(4, 5), num_actions=2, num_planes=512, value_support_size=31,
                                 reward_support_size=31, hidden_dim=64), 'ckpt_cartpole.npz', 16),
    'board_small': ('board', dict(input_shape=(5, 5, 5), num_actions=26, num_res_blocks=2, num_planes=32), None, 12),
}


def build(name):
    kind, kw, ckpt, B = CASES[name]
    cls = mz.MuZeroMLPNet if kind == 'mlp' else mz.MuZeroBoardGameNet
    torch.manual_seed(21)
    net = cls(**kw)
    if ckpt:
        net.load_state_dict({k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLDEN, ckpt)).items()})
    return net.train(), B


@pytest.mark.parametrize('name', list(CASES))
def test_calc_loss_matches_reference_recording(name):
    """tests/golden/train_golden.npz holds what the reference's pipeline.calc_loss produced."""
    z = np.load(os.path.join(GOLDEN, 'train_golden.npz'))
    net, B = build(name)
    tr, w = synthetic_transitions(net, B, 5, seed=77)
    loss, pri = calc_loss(net, 'cpu', tr, torch.from_numpy(w))
    loss.backward()
    assert abs(loss.item() - float(z[f'{name}_loss'])) <= 1e-5 * max(1.0, abs(loss.item()))
    np.testing.assert_allclose(pri, z[f'{name}_priorities'], rtol=1e-5, atol=1e-6)
    for k, p in net.named_parameters():
        g = p.grad.detach().reshape(-1).double()
        got = np.concatenate([[g.sum().item(), g.abs().sum().item()], g[:8].numpy()])
        np.testing.assert_allclose(got, z[f'{name}_grad_{k}'], rtol=1e-4, atol=1e-6, err_msg=k)


def test_stacked_heads_equal_the_reference_loop(monkeypatch):
    """training._calc_loss_stacked (dynamics chain, then the prediction tower, then every head once over the T calls'
    stacked inputs) against the loop of pipeline.py:579-600 on the same network: recorded loss, gradients and every
    BatchNorm buffer (per-call batch statistics, running statistics updated in call order)."""
    import copy
    z = np.load(os.path.join(GOLDEN, 'train_golden.npz'))
    net, B = build('board_small')
    twin = copy.deepcopy(net)
    tr, w = synthetic_transitions(net, B, 5, seed=77)
    monkeypatch.setenv('MZ_TRAIN_BATCHED_HEADS', '0')
    loss_a, pri_a = calc_loss(net, 'cpu', tr, torch.from_numpy(w))
    loss_a.backward()
    monkeypatch.setenv('MZ_TRAIN_BATCHED_HEADS', '2')
    loss_b, pri_b = calc_loss(twin, 'cpu', tr, torch.from_numpy(w))
    loss_b.backward()
    assert abs(loss_b.item() - float(z['board_small_loss'])) <= 1e-5 * max(1.0, abs(loss_b.item()))
    assert abs(loss_a.item() - loss_b.item()) <= 1e-6 * max(1.0, abs(loss_a.item()))
    np.testing.assert_allclose(pri_b, pri_a, rtol=1e-6, atol=1e-7)
    for (k, p), (_, q) in zip(net.named_parameters(), twin.named_parameters()):
        scale = float(p.grad.abs().max()) + 1e-12
        assert float((p.grad - q.grad).abs().max()) <= 2e-5 * scale, k
    for (k, a), (_, b) in zip(net.named_buffers(), twin.named_buffers()):
        torch.testing.assert_close(b.float(), a.float(), rtol=1e-5, atol=1e-7, msg=k)


def test_util_known_answer_vectors():
    """The reference's own KAT (tests/util_test.py:25-47): 3.7 -> 0.3/0.7 on bins 3,4; 2.3 -> 0.7/0.3 on bins 2,3."""
    out = transform_to_2hot(torch.tensor([[3.7], [2.3]]), -5, 5, 11)
    want = torch.zeros(2, 1, 11)
    want[0, 0, 8], want[0, 0, 9] = 0.3, 0.7
    want[1, 0, 7], want[1, 0, 8] = 0.7, 0.3
    torch.testing.assert_close(out, want, rtol=1e-4, atol=1e-4)
    x = torch.linspace(-30, 30, 101)
    torch.testing.assert_close(signed_parabolic(signed_hyperbolic(x)), x, rtol=2e-3, atol=2e-3)
    p = scalar_to_categorical_probabilities(torch.tensor([[0.0, 1.5, -7.0]]), 31)
    assert p.shape == (1, 3, 31) and torch.allclose(p.sum(-1), torch.ones(1, 3), atol=1e-5)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _dp_worker(rank, world, port, name, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    net, B = build(name)
    cfg = mz.make_tictactoe_config(use_tensorboard=False)
    learner = DataParallelLearner(net, cfg, 'cpu')
    tr, w = synthetic_transitions(net, B, 5, seed=77)
    lo, hi = rank * B // world, (rank + 1) * B // world
    shard = Transition(*[x[lo:hi] for x in tr])
    loss, pri = learner.step(shard, w[lo:hi])
    grads = learner.flat_grad.clone()
    params = torch.cat([p.detach().reshape(-1) for p in net.parameters()] +
                       [b.detach().reshape(-1).float() for b in net.buffers()])     # BatchNorm statistics included
    gathered = [torch.zeros_like(params) for _ in range(world)]
    dist.all_gather(gathered, params)
    if rank == 0:
        out.put((loss, grads.numpy(), [g.numpy() for g in gathered]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('name', ['tictactoe_mlp', 'board_small'])
def test_data_parallel_step_world2_gloo(name):
    ctx = mp.get_context('spawn')
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, name, out)) for r in range(2)]
    for p in procs:
        p.start()
    loss0, grads_dp, params = out.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # weights AND BatchNorm running statistics identical on both ranks after the step
    np.testing.assert_array_equal(params[0], params[1])
    # the all-reduced gradient is the mean of the per-shard reference-semantics gradients
    net, B = build(name)
    tr, w = synthetic_transitions(net, B, 5, seed=77)
    want = None
    for r in range(2):
        net_r, _ = build(name)
        shard = Transition(*[x[r * B // 2:(r + 1) * B // 2] for x in tr])
        loss, _ = calc_loss(net_r, 'cpu', shard, torch.from_numpy(w[r * B // 2:(r + 1) * B // 2]))
        loss.backward()
        g = torch.cat([p.grad.reshape(-1) for p in net_r.parameters()]).numpy()
        want = g if want is None else want + g
    np.testing.assert_allclose(grads_dp, want / 2, rtol=1e-4, atol=2e-6)   # thread-count dependent conv reductions
    if name == 'tictactoe_mlp':
        # no BatchNorm: DP over two half batches == the single-process full-batch gradient
        full, _ = build(name)
        loss, _ = calc_loss(full, 'cpu', tr, torch.from_numpy(w))
        loss.backward()
        g_full = torch.cat([p.grad.reshape(-1) for p in full.parameters()]).numpy()
        np.testing.assert_allclose(grads_dp, g_full, rtol=1e-4, atol=1e-6)


def test_learner_checkpoint_keys_match_reference():
    net, _ = build('tictactoe_mlp')
    learner = DataParallelLearner(net, mz.make_tictactoe_config(use_tensorboard=False), 'cpu')
    assert sorted(learner.state_dict()) == ['lr_scheduler', 'network', 'optimizer', 'train_steps']


@pytest.mark.gpu
def test_graphed_training_step_equals_the_eager_one():
    """DataParallelLearner replays one CUDA graph per iteration after three eager ones.  Compared step by step with a
    learner that stays eager, both started from the same weights and Adam state at every iteration (free-running
    trajectories drift apart within a few steps even between two eager learners: cuDNN's backward kernels accumulate
    with atomics and Adam turns the noise of a near-zero gradient into a +-lr step)."""
    import copy
    import muzero_b200 as mz
    from muzero_b200.training import DataParallelLearner, synthetic_transitions
    torch.manual_seed(0)
    net_a = mz.MuZeroBoardGameNet((5, 5, 5), 26, 2, 16).cuda()
    net_b = copy.deepcopy(net_a)
    cfg = mz.config.make_gomoku_config(num_training_steps=10, batch_size=16)
    cfg.lr_milestones = [4]                                      # the schedule must reach the replayed graph
    la = DataParallelLearner(net_a, cfg, 'cuda', use_graph=True)
    lb = DataParallelLearner(net_b, cfg, 'cuda', use_graph=False)
    for it in range(8):
        tr, w = synthetic_transitions(net_a, 16, 5, seed=it)
        net_b.load_state_dict(net_a.state_dict())
        for sa, sb in zip(la.optimizer.state.values(), lb.optimizer.state.values()):
            for k in sa:
                sb[k].copy_(sa[k])
        loss_a, pa = la.step(tr, w)
        loss_b, pb = lb.step(tr, w)
        assert abs(loss_a - loss_b) <= 1e-5 * max(1.0, abs(loss_b)), (it, loss_a, loss_b)
        np.testing.assert_allclose(pa, pb, rtol=1e-4, atol=1e-5)
        gmax = float(lb.flat_grad.abs().max())
        assert float((la.flat_grad - lb.flat_grad).abs().max()) <= 1e-4 * gmax
        for (k, a), (_, b) in zip(net_a.state_dict().items(), net_b.state_dict().items()):
            assert float((a.float() - b.float()).abs().max()) <= 5e-5, (it, k)
        assert abs(float(la.optimizer.param_groups[0]['lr']) - float(lb.optimizer.param_groups[0]['lr'])) < 1e-12
    assert la._graph is not None and lb._graph is None
    assert abs(float(la.optimizer.param_groups[0]['lr']) - cfg.lr_init * cfg.lr_decay_rate) < 1e-9


@pytest.mark.gpu
def test_engine_follows_the_graphed_training_steps():
    """A CUDA-graph replay never advances the parameters' version counters; the learner tells the network explicitly
    (MuZeroNet.mark_weights_updated), so inference after graphed steps must run with the NEW weights."""
    import muzero_b200 as mz
    from oracle.network_oracle import OracleNet
    torch.manual_seed(0)
    kw = dict(input_shape=(9, 3, 3), num_actions=10, num_planes=256, value_support_size=1, reward_support_size=1,
              hidden_dim=64)
    net = mz.MuZeroMLPNet(**kw).cuda()
    cfg = mz.make_tictactoe_config(use_tensorboard=False)
    learner = DataParallelLearner(net, cfg, 'cuda', use_graph=True)
    obs = np.random.RandomState(3).randint(0, 2, size=(8, 9, 3, 3)).astype(np.float32)
    net.eval()
    _, pi_before, _ = net.initial_inference_batch(torch.from_numpy(obs).cuda())
    for it in range(8):                                       # 3 eager + 5 replayed iterations
        tr, w = synthetic_transitions(net, 16, 5, seed=it)
        learner.step(tr, w)
    assert learner._graph is not None
    net.eval()
    _, pi_after, v_after = net.initial_inference_batch(torch.from_numpy(obs).cuda())
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    _, pi_ref, v_ref = OracleNet('mlp', sd, 10, 1, 1).initial_batch(obs)
    np.testing.assert_allclose(pi_after.cpu().numpy(), pi_ref.numpy(), rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(v_after.cpu().numpy(), v_ref.numpy().reshape(-1), rtol=2e-4, atol=2e-4)
    assert float((pi_after - pi_before).abs().max()) > 1e-4   # the weights did move
