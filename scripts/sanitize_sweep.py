"""Stateless sweep only (all of its paths) for a quick compute-sanitizer pass: python scripts/sanitize_sweep.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.getcwd())
import gridworld_b200 as gw  # noqa: E402

n = 200
rng = np.random.RandomState(3)
rnd = lambda rho: np.where(rng.random_sample((n, 9, 11, 11)) < rho, rng.randint(1, 7, (n, 9, 11, 11)), 0).astype(np.int8)  # noqa: E731
for rho in (0.02, 0.3):
    T, G = rnd(rho), rnd(rho)
    G[::2] = np.rot90(T[::2], k=1, axes=(2, 3))
    gw.max_intersection(G, T, return_argmax=True)
    gw.max_intersection(G, T, T)
    gw.max_intersection(G, T, invariant=False)
    G[::5, 3, 4, 5] = 100
    T[1::5, 2, 1, 7] = -100
    gw.max_intersection(G, T, return_argmax=True)
    T[:, 5:] = 0
    T[:, :, :3] = 0
    gw.max_intersection(G, T, return_argmax=True)
print('done')
