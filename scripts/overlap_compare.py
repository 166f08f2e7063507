#!/usr/bin/env python
"""Overlapped step (bits 128 = forced) vs the two-launch / fused forms (bits 64) at several batch sizes, steady state,
obs['grid'] rewritten in full. usage: python scripts/overlap_compare.py [n ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402
import gridworld_b200 as gw  # noqa: E402

dev = torch.device('cuda', 0)
P = 16
for n in [int(a) for a in sys.argv[1:]] or [8192, 16384, 32768, 65536, 262144, 1048576]:
    T = B.synth_targets(n, 0.2, 'compact', 1235).to(dev)
    gen = torch.Generator(device=dev).manual_seed(1)
    acts = torch.randint(0, 18, (P, n), device=dev, dtype=torch.int32, generator=gen)
    out = {'n': n}
    for label, bits in (('overlap', 128), ('plain', 64)):
        for mode in ('every_step', 'on_change'):
            env = gw.make_vec(n, device=dev, max_steps=250, reward_mode=mode, _sched_bits=bits)
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
            out[f'{label}_{mode}'] = round(e0.elapsed_time(e1) / K, 5)
            del env
    print(json.dumps(out), flush=True)
