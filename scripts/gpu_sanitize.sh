#!/bin/bash
# compute-sanitizer hygiene: memcheck + racecheck + initcheck on a small walking run
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
import gridworld_b200 as gw
n = 96
rng = np.random.RandomState(0)
T = np.where(rng.random_sample((n, 9, 11, 11)) < 0.15, rng.randint(1, 7, (n, 9, 11, 11)), 0).astype(np.int8); T[:, 5:] = 0
S = np.where(rng.random_sample((n, 9, 11, 11)) < 0.03, rng.randint(1, 7, (n, 9, 11, 11)), 0).astype(np.int8); S[:, 2:] = 0; S[:, :, 5, 5] = 0
for mode in ('every_step', 'on_change'):
    env = gw.make_vec(n, max_steps=10, reward_mode=mode)
    env.set_tasks(T, S)
    for t in range(25):
        env.step(torch.randint(0, 18, (n,), device='cuda', dtype=torch.int32))
G = np.where(rng.random_sample((n, 9, 11, 11)) < 0.2, rng.randint(1, 7, (n, 9, 11, 11)), 0).astype(np.int8)
gw.max_intersection(G, T, return_argmax=True); gw.max_intersection(G, T, T)
fenv = gw.make_vec(n, action_space='flying', target_in_obs=True); fenv.set_tasks(T, S)
fenv.step({'movement': torch.rand(n, 3, device='cuda'), 'camera': torch.rand(n, 2, device='cuda'),
           'inventory': torch.randint(0, 7, (n,), device='cuda', dtype=torch.int32),
           'placement': torch.randint(0, 3, (n,), device='cuda', dtype=torch.int32)})
torch.cuda.synchronize(); print('sanitize workload done')
PY
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --kernel-regex kns=iglu python /tmp/san.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|done" gpurun_out/sanitizer_$tool.log | tail -3
done
