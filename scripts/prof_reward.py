"""Workload for an ncu capture of the stateless sweep: python scripts/prof_reward.py compact|full|goals rho"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridworld_b200 as gw  # noqa: E402
from bench import synth_targets  # noqa: E402

n = 200000
support, rho = sys.argv[1], float(sys.argv[2])
if support == 'goals':   # bench.py's fourth sweep: the 156 goals.pkl structures tiled, grid = rotated random half
    goals = torch.as_tensor(np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'goals.npz'))['targets'])
    T = goals.to(torch.int8)[torch.arange(n) % goals.shape[0]].cuda()
    keep = torch.rand(T.shape, device='cuda') < 0.5
    G = torch.rot90(torch.where(keep, T, torch.zeros((), dtype=torch.int8, device='cuda')), k=1, dims=(2, 3)).contiguous()
else:
    T = synth_targets(n, rho, support, 1237).cuda()
    G = synth_targets(n, rho, support, 1238).cuda()
    G[:n // 2] = torch.rot90(T[:n // 2], k=1, dims=(2, 3))
for _ in range(3):
    gw.max_intersection(G, T)
torch.cuda.synchronize()
