"""Where the e2e (host-buffer) step time goes: C call (enqueue + kernels + sync) vs Python around it."""
import ctypes as C
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridworld_b200 as gw  # noqa: E402
from bench import synth_targets  # noqa: E402

n = 65536
env = gw.make_vec(n, action_space='walking', max_steps=250, reward_mode='every_step')
env.set_tasks(synth_targets(n, 0.2, 'compact', 1235).cuda())
acts = torch.randint(0, 18, (64, n), dtype=torch.int32).pin_memory()
r = torch.zeros(n, dtype=torch.float64).pin_memory()
d = torch.zeros(n, dtype=torch.uint8).pin_memory()
for t in range(300):
    env.step_host(acts[t % 64], r, d)
K = 1000
t0 = time.perf_counter()
for t in range(K):
    env.step_host(acts[t % 64], r, d)
total = (time.perf_counter() - t0) / K
# the bare C call with pre-marshalled arguments
L, ctx, st, ob = env._L, env._ctx, env._st_ref, env._ob_ref
ptrs = [acts[t].data_ptr() for t in range(64)]
pr, pd, adev = r.data_ptr(), d.data_ptr(), env._actions_ptr
stream = env._stream()
t0 = time.perf_counter()
for t in range(K):
    L.iglu_step_walking_host(ctx, st, ptrs[t % 64], adev, ob, pr, pd, stream)
bare = (time.perf_counter() - t0) / K
# device-only step time for reference
dacts = acts.cuda()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(K):
    env.step(dacts[t % 64])
e1.record()
torch.cuda.synchronize()
print(json.dumps({'step_host_us': total * 1e6, 'bare_c_call_us': bare * 1e6, 'device_step_us': e0.elapsed_time(e1) / K * 1e3}))
