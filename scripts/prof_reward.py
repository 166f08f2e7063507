import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridworld_b200 as gw
from bench import synth_targets
n = 200000
support, rho = sys.argv[1], float(sys.argv[2])
T = synth_targets(n, rho, support, 1237).cuda()
G = synth_targets(n, rho, support, 1238).cuda()
G[:n // 2] = torch.rot90(T[:n // 2], k=1, dims=(2, 3))
for _ in range(3):
    gw.max_intersection(G, T)
torch.cuda.synchronize()
