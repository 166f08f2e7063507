#!/usr/bin/env python
"""BASELINE config 5 in miniature: a torch policy that consumes the device-resident observations IN PLACE
(no host round trip, no int32 grid copy) and feeds its actions straight back into the batched env.

    python examples/policy_loop.py [--envs 1048576] [--steps 200]

The policy is a random-weight MLP over (int8 grid view, agentPos, inventory) -- the point is the data path:
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
    w1 = torch.randn(1089 + 5 + 6, 64, device='cuda', dtype=torch.bfloat16) * 0.05
    w2 = torch.randn(64, 18, device='cuda', dtype=torch.bfloat16) * 0.5

    def policy(o):
        x = torch.cat([o['grid'].reshape(n, -1).to(torch.bfloat16), o['agentPos'].to(torch.bfloat16),
                       o['inventory'].to(torch.bfloat16) * 0.05], dim=1)
        logits = torch.relu(x @ w1) @ w2
        return torch.distributions.Categorical(logits=logits.float()).sample().to(torch.int32)

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
