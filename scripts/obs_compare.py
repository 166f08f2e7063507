#!/usr/bin/env python
"""Tiled vs group-wise observe launch at several batch sizes (steady state, every-step reward).
usage: python scripts/obs_compare.py [n ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402
import gridworld_b200 as gw  # noqa: E402

dev = torch.device('cuda', 0)
P = 16
for n in [int(a) for a in sys.argv[1:]] or [65536, 262144, 1048576]:
    T = B.synth_targets(n, 0.2, 'compact', 1235).to(dev)
    gen = torch.Generator(device=dev).manual_seed(1)
    acts = torch.randint(0, 18, (P, n), device=dev, dtype=torch.int32, generator=gen)
    out = {'n': n}
    for label, bits in (('tiled', 0), ('groups', 32)):
        for name, kw in (('every_full', dict(reward_mode='every_step', grid_obs=True)),
                         ('every_inplace', dict(reward_mode='every_step', grid_obs=False)),
                         ('change_full_split', dict(reward_mode='on_change', grid_obs=True))):
            b = bits | (4 if name.endswith('split') else 0)
            env = gw.make_vec(n, device=dev, max_steps=250, _sched_bits=b, **kw)
            env.set_tasks(T)
            B.preroll(env, acts, n, 7)
            K = max(20, min(400, (65536 * 400) // n))
            for t in range(5):
                env.step(acts[t % P])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for t in range(K):
                env.step(acts[t % P])
            e1.record()
            torch.cuda.synchronize()
            out[f'{label}_{name}'] = round(e0.elapsed_time(e1) / K, 5)
            del env
    print(json.dumps(out), flush=True)
