#!/bin/bash
# Copies the captures of scripts/gpu_r02_profiles.sh from gpurun_out/ into profiles/ (tracked) and derives the two small
# JSON files bench.py quotes (step_kernel_traffic.json, max_intersection_ncu.json).
set -e
cd "$(dirname "$0")/.."
for f in launches.csv launches_summary.txt step_kernels_ncu_summary.txt widen_kernel_ncu_lines.txt physics_kernel_ncu_lines.txt \
         commit_kernel_ncu_lines.txt max_intersection_compact_ncu_summary.txt max_intersection_compact_ncu_lines.txt \
         max_intersection_goals_ncu_summary.txt max_intersection_goals_ncu_lines.txt sass_tma.txt overlap_trace.jsonl; do
  cp gpurun_out/r02_$f profiles/r02_$f
done
python - <<'PY'
import json, re
def kernels(path):
    out, cur = {}, None
    for l in open(path):
        if l.startswith('KERNEL'):
            cur = l.split('KERNEL', 1)[1].strip().split('(')[0].replace('void ', '').strip()
            out[cur] = {}
        else:
            m = re.match(r'\s+(\S+)\s+([\d.,]+)\s*(\S*)', l)
            if m and cur:
                out[cur][m.group(1)] = (float(m.group(2).replace(',', '')), m.group(3))
    return out
k = kernels('profiles/r02_step_kernels_ncu_summary.txt')
def mb(d, key):
    v, u = d[key]
    return int(v * {'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1, 'Gbyte': 1e9}[u])
tot, rec = 0, {}
for name, d in k.items():
    r, w = mb(d, 'dram__bytes_read.sum'), mb(d, 'dram__bytes_write.sum')
    rec[name] = {'dram_bytes_read': r, 'dram_bytes_write': w, 'duration_us': d['gpu__time_duration.sum'][0],
                 'warp_instructions': int(d['smsp__inst_executed.sum'][0])}
    tot += r + w
json.dump({'kernels': 'iglu::k_widen (obs rewrite + pre-step evaluations) || iglu::k_step<0,1> (physics), then iglu::k_commit '
                      '= one step', 'source': 'ncu --set full --clock-control none, steady state of the config-2 workload '
                      '(kernels serialised by the profiler): profiles/r02_step_kernels_ncu_summary.txt',
           'per_kernel': rec, 'dram_bytes_per_launch': tot,
           'note': 'isolated launches: part of the 289 MB of observation writes is still dirty in the 126 MB L2 when k_widen '
                   'ends (it is written back during the next launches); algorithmic bytes per step are 369,229,824'},
          open('profiles/step_kernel_traffic.json', 'w'), indent=1)
mi = {}
for wl, f in (('compact rho=0.20', 'compact'), ('goals.pkl structures tiled; grid = rotated random half of the target', 'goals')):
    d = list(kernels(f'profiles/r02_max_intersection_{f}_ncu_summary.txt').values())[0]
    mi[wl] = {'alu_pipe_pct': d['sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active'][0],
              'issue_active_pct': d['smsp__issue_active.avg.pct_of_peak_sustained_active'][0],
              'source': f'profiles/r02_max_intersection_{f}_ncu_summary.txt'}
json.dump(mi, open('profiles/max_intersection_ncu.json', 'w'), indent=1)
print(json.dumps(rec, indent=1)); print(json.dumps(mi, indent=1))
PY
