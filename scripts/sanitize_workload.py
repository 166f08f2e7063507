"""Small workload that touches every kernel / path of libiglu_b200.so, run under compute-sanitizer."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.getcwd())
import gridworld_b200 as gw  # noqa: E402

n = 96
rng = np.random.RandomState(0)
rnd = lambda rho: np.where(rng.random_sample((n, 9, 11, 11)) < rho, rng.randint(1, 7, (n, 9, 11, 11)), 0).astype(np.int8)  # noqa: E731
T = rnd(0.15)
T[:, 5:] = 0
T[::3, :, 0, :] = 1  # a third of the targets span the zone (single-shift path)
T[::3, :, 10, :] = 2
T[::3, :, :, 0] = 3
T[::3, :, :, 10] = 4
S = rnd(0.03)
S[:, 2:] = 0
S[:, :, 5, 5] = 0
acts = lambda: torch.randint(0, 18, (n,), device='cuda', dtype=torch.int32)  # noqa: E731
for mode in ('every_step', 'on_change'):
    for obs in (True, 'incremental', False):
        env = gw.make_vec(n, max_steps=10, reward_mode=mode, grid_obs=obs)
        env.set_tasks(T, S)
        for t in range(25):
            env.step(acts())
# task generator: resets and auto-resets draw from the pool
env = gw.make_vec(n, max_steps=6, reward_mode='every_step', target_in_obs=True)
env.set_task_pool(T[:17], S[:17], seed=5)
for t in range(20):
    env.step(acts())
env.reset((np.arange(n) % 2).astype(np.uint8))
# host-buffer step, pinned (zero-copy) and pageable (staged)
for pin in (True, False):
    mk = (lambda t: t.pin_memory()) if pin else (lambda t: t)
    a, r, d = mk(torch.zeros(n, dtype=torch.int32)), mk(torch.zeros(n, dtype=torch.float64)), mk(torch.zeros(n, dtype=torch.uint8))
    for t in range(5):
        a.copy_(torch.randint(0, 18, (n,), dtype=torch.int32))
        env.step_host(a, r, d)
# the async form on its own stream + wait, the single-env snapshot, pre-built worlds, env rows evaluated in place
st = torch.cuda.Stream()
for t in range(4):
    a.copy_(torch.randint(0, 18, (n,), dtype=torch.int32))
    with torch.cuda.stream(st):
        env.step_host(a, r, d, wait=False)
    env.wait()
env.fetch_env(3)
env2 = gw.make_vec(n, max_steps=50, reward_mode='every_step')
env2.set_tasks(T, S)
env2.set_grids(np.where(rng.random_sample(T.shape) < 0.5, T, S).astype(np.int8))
for t in range(6):
    env2.step(acts())
gw.max_intersection(env2._grid, env2._target, return_argmax=True)
one = gw.make('IGLUGridworld-v0', render=False, vector_state=True, size_reward=True)
one.set_task(gw.Task('', T[1].astype(np.int32), starting_grid=[(1, -1, 1, 3)]))
for t in range(4):
    one.step(17 if t % 2 else 6)
G = rnd(0.2)
gw.max_intersection(G, T, return_argmax=True)
gw.max_intersection(G, T, T)
gw.max_intersection(G, T, invariant=False)
# crowded zone-spanning pairs: packed-column comparison, and the byte-wise one for rows with stray values
Tf, Gf = rnd(0.3), rnd(0.3)
Gf[::2] = np.rot90(Tf[::2], k=1, axes=(2, 3))
gw.max_intersection(Gf, Tf, return_argmax=True)
Gf[::5, 3, 4, 5] = 100
Tf[1::5, 2, 1, 7] = -100
gw.max_intersection(Gf, Tf, return_argmax=True)
gw.max_intersection(Gf, Tf, invariant=False)
fenv = gw.make_vec(n, action_space='flying', target_in_obs=True)
fenv.set_tasks(T, S)
for t in range(5):
    fenv.step({'movement': torch.rand(n, 3, device='cuda'), 'camera': torch.rand(n, 2, device='cuda'),
               'inventory': torch.randint(0, 7, (n,), device='cuda', dtype=torch.int32),
               'placement': torch.randint(0, 3, (n,), device='cuda', dtype=torch.int32)})
# round 2: the overlapped step (forced for this small batch), the group-wise observe launch, staged tasks, device-made
# task pools and the pov ray caster
for mode in ('every_step', 'on_change'):
    for bits in (128, 32):
        env = gw.make_vec(n, max_steps=8, reward_mode=mode, grid_obs=(True if bits == 128 else False), _sched_bits=bits)
        env.set_task_pool(T[:17], S[:17], seed=9)
        for t in range(20):
            env.step(acts())
renv = gw.make_vec(n, max_steps=8, render=True, render_size=(32, 24), texture=rng.randint(0, 256, (128, 128, 4)).astype(np.uint8),
                   _sched_bits=128)
renv.set_random_task_pool(23, max_blocks=5, height_levels=2, max_dist=2, num_colors=3, task_seed=4)
for t in range(6):
    renv.step(acts())
torch.cuda.synchronize()
print('sanitize workload done')
