"""Error of the tcgen05 recurrent MLP path vs the fp32 torch restatement, per checkpoint (run on a GPU box)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from test_network_gpu import MLPS, load_mlp  # noqa: E402

for name in MLPS:
    net, onet, kw = load_mlp(name)
    for batch in (1, 100, 4096):
        gen = np.random.RandomState(batch)
        obs = gen.standard_normal((batch,) + kw['input_shape']).astype(np.float32)
        h_ref, _, _ = onet.initial_batch(obs)
        act = gen.randint(0, kw['num_actions'], size=batch)
        slots_in = net.hidden_from_reference(h_ref.cuda())
        out, reward, pi2, value2 = net.recurrent_inference_batch(slots_in, torch.from_numpy(act).cuda())
        h2_ref, r_ref, pi2_ref, v2_ref = onet.recurrent_batch(h_ref, act)
        def err(a, b):
            a, b = a.cpu().numpy().astype(np.float64), b.numpy().astype(np.float64)
            return f'max|err| {np.abs(a - b).max():.3g} (ref |max| {np.abs(b).max():.3g})'
        print(name, batch, 'h', err(net.hidden_to_reference(out), h2_ref), '| r', err(reward, r_ref), '| v', err(value2, v2_ref),
              '| pi', err(pi2, pi2_ref), flush=True)
