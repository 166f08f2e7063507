"""Observe-launch scheduling sweep (static deal vs dynamic groups of g envs), BASELINE config 2 workload.
usage: python scripts/sched_sweep.py [envs] -> one JSON line per variant (ms per step over 1000 steps after 300)."""
import json
import os
import sys

import torch

sys.path.insert(0, '.')
import gridworld_b200 as gw  # noqa: E402
from bench import synth_targets  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = synth_targets(n, 0.2, 'compact', 1235).cuda()
acts = torch.randint(0, 18, (64, n), device='cuda', dtype=torch.int32)
variants = [('static', 8)] + [(f'dynamic g={g}', g << 8) for g in [int(x) for x in os.environ.get('G', '4,6,8,12,16').split(',')]]
for name, bits in variants:
    for rm, obs in (('every_step', True), ('every_step', 'incremental')):
        env = gw.make_vec(n, max_steps=250, reward_mode=rm, grid_obs=obs, _sched_bits=bits)
        env.set_tasks(T)
        for t in range(300):
            env.step(acts[t % 64])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 1000 if n <= 65536 else 100
        e0.record()
        for t in range(K):
            env.step(acts[t % 64])
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({'variant': name, 'reward_mode': rm, 'grid_obs': obs, 'envs': n,
                          'ms_per_step': e0.elapsed_time(e1) / K}), flush=True)
        del env
