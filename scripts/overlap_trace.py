#!/usr/bin/env python
"""Timeline of the overlapped step (k_widen || physics, then k_commit) at steady state from the library's tracing events
(IGLU_B200_TRACE=1): end of the widen + physics pair and start / end of the commit launch, relative to the start of the
widen launch. (The round-2 two-stream prototype could time the two concurrent launches separately:
profiles/r02_overlap_trace_two_streams.jsonl.) usage: python scripts/overlap_trace.py [n] [sched_bits]"""
import ctypes as C
import json
import os
import sys

os.environ['IGLU_B200_TRACE'] = '1'
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402
import gridworld_b200 as gw  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device('cuda', 0)
P = 16
T = B.synth_targets(n, 0.2, 'compact', 1235).to(dev)
gen = torch.Generator(device=dev).manual_seed(1)
acts = torch.randint(0, 18, (P, n), device=dev, dtype=torch.int32, generator=gen)
for mode in ('every_step', 'on_change'):
    env = gw.make_vec(n, device=dev, max_steps=250, reward_mode=mode, _sched_bits=bits)
    env.set_tasks(T)
    B.preroll(env, acts, n, 7)
    rows = []
    buf = (C.c_float * 5)()
    for t in range(60):
        env.step(acts[t % P])
        if t >= 20:
            gw._lib.check(env._L.iglu_debug_overlap_trace(env._ctx, buf))
            rows.append(list(buf))
    r = np.median(np.array(rows), axis=0) * 1e3
    print(json.dumps({'n': n, 'mode': mode, 'bits': bits, 'widen_and_physics_end_us': round(float(r[2]), 1),
                      'commit_start_us': round(float(r[3]), 1), 'commit_end_us': round(float(r[4]), 1)}), flush=True)
    del env
