#!/usr/bin/env python
"""Stateless sweep on FULL-support pairs (targets span the zone: one shift per rotation, the byte-compare path):
evals/s and the fraction of the measured HBM bandwidth (2 x 1089 + 4 B per pair). usage: python scripts/mi_full_support.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402
import gridworld_b200 as gw  # noqa: E402

dev = torch.device('cuda', 0)
n = 1_000_000
peak, _ = B.hbm_peak()
for rho in (0.05, 0.2, 0.5):
    T = B.synth_targets(n, rho, 'full', 1237, device=dev)
    G = B.synth_targets(n, rho, 'full', 1238, device=dev)
    G[:n // 2] = torch.rot90(T[:n // 2], k=1, dims=(2, 3))
    gw.max_intersection(G, T)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        gw.max_intersection(G, T)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    ev = n / (ms * 1e-3)
    print(json.dumps({'support': 'full', 'rho': rho, 'ms': round(ms, 4), 'evals_per_sec': ev,
                      'hbm_gbs': ev * (2 * 1089 + 4) / 1e9, 'hbm_frac': ev * (2 * 1089 + 4) / 1e9 / peak}), flush=True)
    del T, G
