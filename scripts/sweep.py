#!/usr/bin/env python
"""Throughput sweep (not the bench contract): step-path variants and the stateless reward kernel
(BASELINE configs 2-4) timed with CUDA events. Prints one JSON object per line; copy into profiles/."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridworld_b200 as gw  # noqa: E402
from bench import synth_targets  # noqa: E402


def time_steps(env, step_fn, K=300, W=20):
    for t in range(W):
        step_fn(t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(K):
        step_fn(W + t)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K


def walking(n, support, rho, reward_mode, inplace, K=300):
    env = gw.make_vec(n, action_space='walking', max_steps=250, reward_mode=reward_mode,
                      grid_obs='incremental' if inplace == 2 else (not inplace),
                      alternate_phases=not os.environ.get('IGLU_NO_ALT'),
                      split_launch='force' if os.environ.get('IGLU_FORCE_SPLIT') else (not os.environ.get('IGLU_NO_SPLIT')))
    env.set_tasks(synth_targets(n, rho, support, 1235).cuda())
    acts = torch.randint(0, 18, (64, n), device='cuda', dtype=torch.int32)
    ms = time_steps(env, lambda t: env.step(acts[t % 64]), K=K)
    return {'kind': 'walking', 'envs': n, 'support': support, 'rho': rho, 'reward_mode': reward_mode,
            'inplace': inplace, 'ms_per_step': ms, 'env_steps_per_sec': n / ms * 1e3}


def flying(n, K=200):
    env = gw.make_vec(n, action_space='flying', max_steps=250, target_in_obs=True, reward_mode='on_change')
    env.set_tasks(synth_targets(n, 0.2, 'compact', 1236).cuda())
    mv = torch.rand(16, n, 3, device='cuda') * 2 - 1
    cam = torch.rand(16, n, 2, device='cuda') * 10 - 5
    inv = torch.randint(0, 7, (16, n), device='cuda', dtype=torch.int32)
    pl = torch.randint(0, 3, (16, n), device='cuda', dtype=torch.int32)
    ms = time_steps(env, lambda t: env.step({'movement': mv[t % 16], 'camera': cam[t % 16],
                                             'inventory': inv[t % 16], 'placement': pl[t % 16]}), K=K)
    return {'kind': 'flying+target_in_obs', 'envs': n, 'ms_per_step': ms, 'env_steps_per_sec': n / ms * 1e3}


def p_eval_mean(T, limit=2000):
    """SURVEY 8(d) algorithmic work per evaluation: P_eval = sum_k S_k * NR_k popcount-ANDs, with S_k the
    admissible shifts of rotation k and NR_k the non-empty (class, y, x) rows of the rotated target."""
    T = T[:limit].cpu().numpy()
    tot = 0.0
    for t in T:
        nz = np.argwhere(t != 0)
        if len(nz) == 0:
            continue
        xmin, xmax, zmin, zmax = nz[:, 1].min(), nz[:, 1].max(), nz[:, 2].min(), nz[:, 2].max()
        s_even = (xmin - (xmax - 10) + 1) * (zmin - (zmax - 10) + 1)  # same count for all four rotations
        cls = t[nz[:, 0], nz[:, 1], nz[:, 2]]
        nr_x = len({(c, y, x) for c, (y, x, z) in zip(cls, nz)})   # rotations 0 and 2: rows along x
        nr_z = len({(c, y, z) for c, (y, x, z) in zip(cls, nz)})   # rotations 1 and 3: rows along z
        tot += s_even * (2 * nr_x + 2 * nr_z)
    return tot / len(T)


_INT_PEAK = None


def int_peak():
    """thread-ops/s of the POPC(+IADD), LOP3, IADD3, IMAD chains measured by csrc/iglu_microbench.cu."""
    global _INT_PEAK
    if _INT_PEAK is None:
        import ctypes as C
        from gridworld_b200 import build as B
        lib = C.CDLL(B.build_microbench())
        out = (C.c_double * 4)()
        lib.iglu_mb_int_peak.argtypes = [C.POINTER(C.c_double), C.c_void_p]
        torch.cuda.synchronize()
        assert lib.iglu_mb_int_peak(out, None) == 0
        _INT_PEAK = dict(zip(('popc', 'lop3', 'iadd3', 'imad'), list(out)))
    return _INT_PEAK


def reward_only(n, support, rho, reps=5):
    T = synth_targets(n, rho, support, 1237).cuda()
    Gd = synth_targets(n, rho, support, 1238).cuda()
    h = n // 2  # half the grids: rotated copy of the target
    Gd[:h] = torch.rot90(T[:h], k=1, dims=(2, 3))
    gw.max_intersection(Gd, T)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        gw.max_intersection(Gd, T)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    pe = p_eval_mean(T)
    pk = int_peak()
    evals = n / ms * 1e3
    return {'kind': 'max_intersection', 'pairs': n, 'support': support, 'rho': rho, 'ms': ms,
            'evals_per_sec': evals, 'p_eval_mean': pe, 'popcount_ands_per_sec': evals * pe,
            'popc_peak_thread_ops_per_sec': pk['popc'], 'frac_of_popc_peak': evals * pe / pk['popc'],
            'hbm_bytes_per_eval': 2 * 1089 + 4, 'hbm_gbs': evals * (2 * 1089 + 4) / 1e9, 'int_peaks': pk}


def goals_tiled(n, reward_mode, K=200):
    """The 156 real goal structures of skills/goals.pkl (tests/golden/goals.npz) tiled over n envs."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    T = torch.as_tensor(np.load(os.path.join(root, 'tests', 'golden', 'goals.npz'))['targets'])
    T = T[torch.arange(n) % T.shape[0]].cuda()
    env = gw.make_vec(n, action_space='walking', max_steps=250, reward_mode=reward_mode)
    env.set_tasks(T)
    acts = torch.randint(0, 18, (64, n), device='cuda', dtype=torch.int32)
    ms = time_steps(env, lambda t: env.step(acts[t % 64]), K=K)
    return {'kind': 'walking, goals.pkl targets tiled', 'envs': n, 'reward_mode': reward_mode, 'ms_per_step': ms,
            'env_steps_per_sec': n / ms * 1e3}


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'one':  # one: <reward_mode> <inplace 0/1> <envs> [steps]
        K = int(sys.argv[5]) if len(sys.argv) > 5 else 30
        print(json.dumps(walking(int(sys.argv[4]), 'compact', 0.2, sys.argv[2], int(sys.argv[3]), K=K)))
        sys.exit(0)
    which = sys.argv[1:] or ['walking', 'flying', 'reward']
    if 'walking' in which:
        for rm in ('every_step', 'on_change'):
            for inplace in (False, True):
                print(json.dumps(walking(65536, 'compact', 0.2, rm, inplace)), flush=True)
        print(json.dumps(walking(65536, 'full', 0.2, 'every_step', False)), flush=True)
        print(json.dumps(walking(1048576, 'compact', 0.2, 'on_change', True, K=50)), flush=True)
        print(json.dumps(walking(1048576, 'compact', 0.2, 'every_step', False, K=30)), flush=True)
    if 'goals' in which:
        for rm in ('every_step', 'on_change'):
            print(json.dumps(goals_tiled(65536, rm)), flush=True)
    if 'flying' in which:
        print(json.dumps(flying(262144)), flush=True)
    if 'reward' in which:
        for rho in (0.05, 0.2, 0.5):
            for support in ('full', 'compact'):
                print(json.dumps(reward_only(1_000_000, support, rho)), flush=True)
