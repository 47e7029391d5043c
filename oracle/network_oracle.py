"""TEST INFRASTRUCTURE ONLY — fp32 PyTorch-CPU restatement of the reference networks.

Functional restatement of ``muzero/network.py`` + ``muzero/util.py`` driven by a
reference ``state_dict`` (same key names).  It is the "plain torch fp32
reference" the floating-point kernels are compared against, and the network
half of the CPU baseline (``bench.py --impl reference``).  Pinned against the
real reference modules by ``tests/golden/make_golden_nets.py`` (bit-identical
on the same machine: same torch ops in the same order).
"""
from __future__ import annotations

from typing import Dict, NamedTuple

import numpy as np
import torch
import torch.nn.functional as F


class NetworkOutputs(NamedTuple):
    hidden_state: np.ndarray
    reward: float
    pi_probs: np.ndarray
    value: float


# --- util.py -----------------------------------------------------------------
def signed_parabolic(x, eps: float = 1e-3):
    """util.py:25-28."""
    z = torch.sqrt(1 + 4 * eps * (eps + 1 + torch.abs(x))) / 2 / eps - 1 / 2 / eps
    return torch.sign(x) * (torch.square(z) - 1)


def signed_hyperbolic(x, eps: float = 1e-3):
    """util.py:20-22."""
    return torch.sign(x) * (torch.sqrt(torch.abs(x) + 1) - 1) + eps * x


def normalize_hidden_state(h):
    """util.py:31-36: min/max over dim 1 only."""
    lo = h.min(dim=1, keepdim=True)[0]
    hi = h.max(dim=1, keepdim=True)[0]
    return (h - lo) / (hi - lo + 1e-8)


def logits_to_transformed_expected_value(logits, support_size: int):
    """util.py:70-93 (+ transform_from_2hot, util.py:62-67)."""
    hi = (support_size - 1) // 2
    probs = torch.softmax(logits, dim=-1)
    support = torch.linspace(-hi, hi, support_size).expand_as(probs)
    x = torch.sum(probs * support, dim=-1, keepdim=True)
    return signed_parabolic(x)


# --- network.py ----------------------------------------------------------------
class OracleNet:
    """kind: 'mlp' | 'board' | 'atari'."""

    def __init__(self, kind: str, sd: Dict[str, torch.Tensor], num_actions: int, value_support: int = 1,
                 reward_support: int = 1, num_res_blocks: int = 0):
        self.kind, self.A = kind, num_actions
        self.sv, self.sr, self.blocks = value_support, reward_support, num_res_blocks
        self.sd = {k: torch.as_tensor(v).to(torch.float32) for k, v in sd.items()}

    # building blocks -----------------------------------------------------------
    def _lin(self, p, x):
        return F.linear(x, self.sd[p + '.weight'], self.sd[p + '.bias'])

    def _mlp2(self, p, x):                                    # Linear-ReLU-Linear (network.py:145-149)
        return self._lin(p + '.2', F.relu(self._lin(p + '.0', x)))

    def _bn(self, p, x):                                      # eval-mode BatchNorm2d
        return F.batch_norm(x, self.sd[p + '.running_mean'], self.sd[p + '.running_var'], self.sd[p + '.weight'],
                            self.sd[p + '.bias'], False, 0.0, 1e-5)

    def _conv(self, p, x, stride=1, pad=1):
        return F.conv2d(x, self.sd[p + '.weight'], None, stride, pad)

    def _block(self, p, x):                                   # ResNetBlock, network.py:273-299
        o = F.relu(self._bn(p + '.conv_block1.1', self._conv(p + '.conv_block1.0', x)))
        o = self._bn(p + '.conv_block2.1', self._conv(p + '.conv_block2.0', o))
        return F.relu(o + x)

    def _tower(self, p, x, n):
        for i in range(n):
            x = self._block(f'{p}.{i}', x)
        return x

    def _head(self, p, x):                                    # 1x1 conv-BN-ReLU-Flatten-Linear
        o = F.relu(self._bn(p + '.1', self._conv(p + '.0', x, 1, 0)))
        return self._lin(p + '.4', o.flatten(1))

    # the three functions ----------------------------------------------------------
    def represent(self, x):
        if self.kind == 'mlp':
            h = self._mlp2('represent_net.net', x.reshape(x.shape[0], -1))
        elif self.kind == 'board':                            # network.py:356-393
            h = F.relu(self._bn('represent_net.conv_block.1', self._conv('represent_net.conv_block.0', x)))
            h = self._tower('represent_net.res_blocks', h, self.blocks)
        else:                                                 # network.py:312-353
            h = F.relu(self._conv('represent_net.conv_1', x, 2, 1))
            h = self._tower('represent_net.res_blocks_1', h, 2)
            h = F.relu(self._conv('represent_net.conv_2', h, 2, 1))
            h = self._tower('represent_net.res_blocks_2', h, 2)
            h = F.avg_pool2d(h, 3, 2, 1)
            h = self._tower('represent_net.res_blocks_3', h, 2)
            h = F.avg_pool2d(h, 3, 2, 1)
        return normalize_hidden_state(h)

    def dynamics(self, h, action):
        b = h.shape[0]
        if self.kind == 'mlp':                                # network.py:182-198
            onehot = torch.zeros((b, self.A), dtype=torch.float32).scatter_(1, action.reshape(b, 1), 1.0)
            hs = self._mlp2('dynamics_net.transition_net', torch.cat([h, onehot], dim=1))
            r = self._mlp2('dynamics_net.reward_net', hs)
        else:                                                 # network.py:432-449
            hh, ww = h.shape[2], h.shape[3]
            # QUIRK C (network.py:440-444): the action arrives as [B, 1], so one_hot is [B, 1, A],
            # repeat_interleave(dim=1) gives [B, h*w, A] and the reshape to [B, A, h, w] does NOT
            # produce constant planes: flat element f of the [A*h*w] block is 1 iff f % A == action.
            planes = F.one_hot(action.reshape(b, 1).long(), self.A).to(torch.float32)
            planes = torch.repeat_interleave(planes, repeats=hh * ww, dim=1).reshape(b, self.A, hh, ww)
            x = torch.cat([h, planes], dim=1)
            hs = F.relu(self._bn('dynamics_net.conv_block.1', self._conv('dynamics_net.conv_block.0', x)))
            hs = self._tower('dynamics_net.res_blocks', hs, self.blocks)
            r = self._head('dynamics_net.reward_head', hs)
        return normalize_hidden_state(hs), r

    def prediction(self, h):
        if self.kind == 'mlp':
            return self._mlp2('prediction_net.policy_net', h), self._mlp2('prediction_net.value_net', h)
        f = self._tower('prediction_net.res_blocks', h, self.blocks)
        return self._head('prediction_net.policy_net', f), self._head('prediction_net.value_net', f)

    # batched inference (tensors) ---------------------------------------------------
    @torch.no_grad()
    def initial_batch(self, x):
        h = self.represent(torch.as_tensor(x, dtype=torch.float32))
        pl, v = self.prediction(h)
        pi = F.softmax(pl, dim=1)
        if self.sv != 1:
            v = logits_to_transformed_expected_value(v, self.sv)
        return h, pi, v.reshape(-1)

    @torch.no_grad()
    def recurrent_batch(self, h, action):
        h, r = self.dynamics(torch.as_tensor(h, dtype=torch.float32), torch.as_tensor(action, dtype=torch.long))
        if self.sr != 1:
            r = logits_to_transformed_expected_value(r, self.sr)
        pl, v = self.prediction(h)
        pi = F.softmax(pl, dim=1)
        if self.sv != 1:
            v = logits_to_transformed_expected_value(v, self.sv)
        return h, r.reshape(-1), pi, v.reshape(-1)

    # the reference's single-item API (network.py:62-111) --------------------------------
    def initial_inference(self, x):
        h, pi, v = self.initial_batch(x)
        return NetworkOutputs(h.squeeze(0).numpy(), 0.0, pi.squeeze(0).numpy(), v.squeeze(0).item())

    def recurrent_inference(self, hidden_state, action):
        h, r, pi, v = self.recurrent_batch(hidden_state, action)
        return NetworkOutputs(h.squeeze(0).numpy(), r.squeeze(0).item(), pi.squeeze(0).numpy(), v.squeeze(0).item())


def randomize_batchnorm(net, seed: int) -> None:
    """Give every BatchNorm2d of a torch module non-trivial, reproducible statistics and affine
    parameters (module traversal order).  tests/golden/make_golden_nets.py used exactly this
    recipe on the reference modules before recording their outputs."""
    g = torch.Generator().manual_seed(seed)
    for m in list(net.modules()):
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) * 0.5 + 0.75)
            m.weight.data.copy_(torch.rand(m.num_features, generator=g) * 0.5 + 0.75)
            m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.1)
