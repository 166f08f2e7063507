"""Where does the per-step time go: CPU enqueue cost of env.step vs GPU time, eager vs CUDA-graph replay."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridworld_b200 as gw
from bench import synth_targets

n = 65536
env = gw.make_vec(n, action_space='walking', max_steps=250, reward_mode='every_step')
env.set_tasks(synth_targets(n, 0.2, 'compact', 1235).cuda())
acts = torch.randint(0, 18, (64, n), device='cuda', dtype=torch.int32)
for t in range(50):
    env.step(acts[t % 64])
torch.cuda.synchronize()
K = 1000
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for t in range(K):
    env.step(acts[t % 64])
e1.record(); t1 = time.perf_counter()
torch.cuda.synchronize()
out = {'eager_gpu_ms_per_step': e0.elapsed_time(e1) / K, 'eager_cpu_enqueue_ms_per_step': (t1 - t0) * 1e3 / K}
# CUDA graph of 64 consecutive steps (static action slices)
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for t in range(3):
        env.step(acts[t])
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    for t in range(64):
        env.step(acts[t])
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
R = 16
e0.record()
for _ in range(R):
    g.replay()
e1.record()
torch.cuda.synchronize()
out['graph_gpu_ms_per_step'] = e0.elapsed_time(e1) / (R * 64)
print(json.dumps(out))
