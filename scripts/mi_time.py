"""Quick timing of the stateless reward sweep: python scripts/mi_time.py [pairs] -> ms per launch for three inputs."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridworld_b200 as gw  # noqa: E402
from bench import synth_targets  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
out = {}
for support, rho in (('compact', 0.2), ('compact', 0.5), ('full', 0.2)):
    T = synth_targets(n, rho, support, 1237).cuda()
    G = synth_targets(n, rho, support, 1238).cuda()
    G[:n // 2] = torch.rot90(T[:n // 2], k=1, dims=(2, 3))
    for _ in range(3):
        gw.max_intersection(G, T)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        gw.max_intersection(G, T)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out[f'{support} {rho}'] = {'ms': ms, 'evals_per_sec': n / ms * 1e3}
print(json.dumps(out))
