#!/usr/bin/env python
"""One JSON line of steady-state timings for the library selected with IGLU_B200_SO (tuning aid):
headline (config 2, desynchronised steady state), dense (every env evaluated), in-place variants, reward sweep.
usage: [IGLU_B200_SO=...] python scripts/tune_r2.py [label] [sched_bits]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402
import gridworld_b200 as gw  # noqa: E402

label = sys.argv[1] if len(sys.argv) > 1 else os.path.basename(os.environ.get('IGLU_B200_SO', 'default'))
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device('cuda', 0)
n, P = 65536, 64
T = B.synth_targets(n, 0.2, 'compact', 1235).to(dev)
gen = torch.Generator(device=dev).manual_seed(1)
acts = torch.randint(0, 18, (P, n), device=dev, dtype=torch.int32, generator=gen)


def timed(env, K=400):
    for t in range(10):
        env.step(acts[t % P])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(K):
        env.step(acts[t % P])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K


out = {'label': label, 'bits': bits}
for name, kw in (('every_full', dict(reward_mode='every_step', grid_obs=True)),
                 ('every_inplace', dict(reward_mode='every_step', grid_obs=False)),
                 ('change_full', dict(reward_mode='on_change', grid_obs=True)),
                 ('change_inplace', dict(reward_mode='on_change', grid_obs=False))):
    env = gw.make_vec(n, device=dev, max_steps=250, _sched_bits=bits, **kw)
    env.set_tasks(T)
    B.preroll(env, acts, n, 7)
    out[name] = round(timed(env), 5)
    del env
env = gw.make_vec(n, device=dev, max_steps=1 << 30, reward_mode='every_step', _sched_bits=bits)
env.set_tasks(T)
built = torch.where(torch.rand(T.shape, device=dev, generator=gen) < 0.5, T, torch.zeros((), dtype=torch.int8, device=dev))
built[:, :3, 5, 5] = 0
env.set_grids(built)
out['dense_full'] = round(timed(env, 200), 5)
del env
for rho in (0.05, 0.2, 0.5):
    Tp = B.synth_targets(1_000_000, rho, 'compact', 1237, device=dev)
    Gp = B.synth_targets(1_000_000, rho, 'compact', 1238, device=dev)
    Gp[:500_000] = torch.rot90(Tp[:500_000], k=1, dims=(2, 3))
    gw.max_intersection(Gp, Tp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        gw.max_intersection(Gp, Tp)
    e1.record()
    torch.cuda.synchronize()
    out[f'mi_{rho}'] = round(1e6 / (e0.elapsed_time(e1) / 5 * 1e-3) / 1e8, 3)   # 1e8 evals/s
    del Tp, Gp
print(json.dumps(out), flush=True)
