#!/usr/bin/env python
"""BASELINE config 5 in miniature: a torch policy that consumes the device-resident observations IN PLACE
(no host round trip, no int32 grid copy) and feeds its actions straight back into the batched env.

    python examples/policy_loop.py [--envs 1048576] [--steps 200]

The policy is a random-weight MLP over (int8 state rows, agentPos, inventory) -- the point is the data path:
env.step() returns tensors that alias persistent device buffers; env.grid is the int8 voxel state itself.
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridworld_b200 as gw  # noqa: E402
from bench import synth_targets  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--envs', type=int, default=262144)
    ap.add_argument('--steps', type=int, default=200)
    args = ap.parse_args()
    n = args.envs
    env = gw.make_vec(n, action_space='walking', reward_mode='on_change', grid_obs=False, max_steps=250)
    env.set_tasks(synth_targets(n, 0.2, 'compact', 99).cuda())
    obs = env.reset()
    torch.manual_seed(0)
    w1 = torch.randint(-8, 9, (1104, 64), device='cuda', dtype=torch.int8)   # K = the padded state row (pad cells are 0)
    w1s = torch.randn(5 + 6, 64, device='cuda', dtype=torch.bfloat16) * 0.05
    w2 = torch.randn(64, 18, device='cuda', dtype=torch.bfloat16) * 0.5

    def policy(o):
        # env.grid_rows is the int8 voxel state where it lies (rows of 1104 bytes): the first layer is an int8 GEMM on it,
        # no widening pass (torch's own int8 -> bf16 conversion of 1 M rows costs 2.8 ms, the int8 GEMM 0.3 ms)
        h = torch._int_mm(env.grid_rows, w1).to(torch.bfloat16) * 0.006
        h = h + torch.cat([o['agentPos'], o['inventory'] * 0.05], dim=1).to(torch.bfloat16) @ w1s
        logits = (torch.relu(h) @ w2).float()
        gumbel = -torch.log(-torch.log(torch.rand_like(logits).clamp_min(1e-20)))
        return (logits + gumbel).argmax(dim=1).to(torch.int32)   # a sample of Categorical(logits)

    ret = torch.zeros(n, dtype=torch.float64, device='cuda')
    for _ in range(5):
        obs, r, d, _ = env.step(policy(obs))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        obs, r, d, _ = env.step(policy(obs))
        ret += r
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f'{n} envs x {args.steps} steps with an on-device policy: {n * args.steps / dt:.3e} env-steps/s '
          f'({1e3 * dt / args.steps:.3f} ms/step, mean return so far {ret.mean().item():.4f})')


if __name__ == '__main__':
    main()
