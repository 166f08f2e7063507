#!/usr/bin/env python
"""bench.py -- env-steps/sec of the IGLU Gridworld step path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] ...                # the CPU arm (oracle port, all host threads)
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # one rank per GPU, envs sharded, no collective

Workload (config.workload): BASELINE.json configs[1] -- walking Discrete(18), 65,536 envs per GPU,
random synthetic target structures (SURVEY 8(d) generator, compact support, density 0.20: ~25 blocks and
~196 admissible (rotation, shift) pairs per task, close to the real goals.pkl mix), uniform random
actions, full maximal_intersection evaluated on EVERY step, int32 obs['grid'] materialised every step,
auto-reset on done. A "step" = one call of the step path over all envs of the rank.

value  = env-steps/s with the action tensors already resident in HBM (CUDA events, max over ranks).
e2e    = the same metric through the host-buffer C-ABI call iglu_step_walking_host: every step takes the
         int32 actions from pinned host memory and delivers reward (f64) + done (u8) into pinned host memory,
         stream synchronised per step. The buffers are pinned, so the kernels move these bytes over PCIe
         themselves (zero-copy reads / writes inside the step kernels) instead of staging them through
         cudaMemcpyAsync before / after the step; `e2e.staged_value` is the same call with the staged copies
         forced. Observations stay on the device for an on-device policy (BASELINE configs[4]).
roofline: HBM. achieved = A_walk * envs / mean step-kernel duration, A_walk = 5,634 algorithmic bytes per
         env-step (4,413 B of observations/results written, 1,089 B grid + 4 B action read,
         64 B state read + 64 B written), peak = MEASURED_PEAKS.json hbm_gbs.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

A_WALK_BYTES = 5634          # SURVEY 8(d) algorithmic bytes per walking env-step, materialised obs
ENVS_PER_GPU = 65536         # BASELINE.json configs[1]
DENSITY = 0.20
MAX_STEPS = 250
FALLBACK_HBM_GBS = 6650.0    # /opt/skills/guides/B200_PROFILING.md fallback


def synth_targets(n, rho, support, seed):
    """target = (u < rho) ? 1 + floor(6 v) : 0, u, v ~ U[0,1) from torch.Generator('cpu')."""
    import torch
    g = torch.Generator('cpu').manual_seed(seed)
    u = torch.rand(n, 9, 11, 11, generator=g)
    v = torch.rand(n, 9, 11, 11, generator=g)
    t = torch.where(u < rho, 1 + torch.floor(6 * v), torch.zeros(())).to(torch.int8)
    if support == 'compact':
        m = torch.zeros(9, 11, 11, dtype=torch.bool)
        m[:5, 3:8, 3:8] = True
        t = torch.where(m, t, torch.zeros((), dtype=torch.int8))
    return t


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:  # noqa: BLE001
            self.ok = False

    NAMES = {0x1: 'gpu_idle', 0x2: 'applications_clocks_setting', 0x4: 'sw_power_cap', 0x8: 'hw_slowdown',
             0x10: 'sync_boost', 0x20: 'sw_thermal_slowdown', 0x40: 'hw_thermal_slowdown',
             0x80: 'hw_power_brake_slowdown', 0x100: 'display_clock_setting'}

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.NAMES.items():
                    if r & bit and name != 'gpu_idle':
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = int(np.median(self.samples)) if self.samples else None
        return {'sm_mhz': med, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(self.samples)}


def physical_gpu_index(local_index):
    vis = os.environ.get('CUDA_VISIBLE_DEVICES')
    if vis:
        try:
            return int(vis.split(',')[local_index])
        except Exception:  # noqa: BLE001
            return local_index
    return local_index


def hbm_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:  # noqa: BLE001
            pass
    return FALLBACK_HBM_GBS, 'fallback (B200_PROFILING.md)'


def recorded_traffic():
    """dram bytes per launch of the step kernel from the committed ncu --set full capture, or None."""
    p = os.path.join(ROOT, 'profiles', 'step_kernel_traffic.json')
    if os.path.exists(p):
        try:
            return json.load(open(p)).get('dram_bytes_per_launch')
        except Exception:  # noqa: BLE001
            return None
    return None


WORKLOAD = (f'configs[1]: walking Discrete(18), {ENVS_PER_GPU} envs/GPU, synthetic compact targets rho={DENSITY} '
            f'(~25 blocks, ~196 admissible rot x shift), uniform random actions, maximal_intersection every step, '
            f'int32 grid obs materialised, auto-reset, max_steps={MAX_STEPS}')


# --------------------------------------------------------------------------------------- CPU arm

def cpu_sample(n_envs, steps, threads, seed=1234 + 1, warmup=0):
    """Times the oracle port (C restatement of the reference step) on `n_envs` envs of the workload."""
    from oracle import oracle as O
    T = synth_targets(n_envs, DENSITY, 'compact', seed).numpy()
    b = O.OracleBatch(n_envs, 'walking', select_and_place=True, max_steps=MAX_STEPS, threads=threads)
    b.set_tasks(T)
    rng = np.random.RandomState(0)
    acts = rng.randint(0, 18, (64, n_envs)).astype(np.int32)
    for t in range(warmup):
        b.step_walking(acts[t % 64], auto_reset=True)
    t0 = time.perf_counter()
    for t in range(steps):
        b.step_walking(acts[t % 64], auto_reset=True)
    dt = time.perf_counter() - t0
    return n_envs * steps / dt, dt


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_envs = 8192
    # size the per-step sample so that steps + warmup finish within a few minutes
    rate, _ = cpu_sample(2048, 20, threads)
    budget_s = 120.0
    per_step = n_envs / rate
    if per_step * (args.steps + args.warmup) > budget_s:
        n_envs = max(256, int(rate * budget_s / (args.steps + args.warmup)) // 256 * 256)
    value, dt = cpu_sample(n_envs, args.steps, threads, warmup=args.warmup)
    line = {
        'impl': 'reference', 'metric': 'env_steps_per_sec', 'value': value, 'unit': 'env-steps/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'envs_per_step_sample': n_envs,
                   'note': 'reference is pure Python (cannot travel to the GPU box); this arm times its C '
                           'restatement (oracle/iglu_oracle.c, pinned bit-exact to the executed reference) '
                           'with one pthread per host core; reward gate as in the reference (on change)'},
        'cpu_baseline': {'value': value, 'unit': 'env-steps/s', 'cores': threads, 'kind': 'port',
                         'sample': f'{n_envs} envs x {args.steps} steps of the workload'},
        'e2e': {'value': value, 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------- GPU arm

def run_b200_arm(args):
    import torch
    from gridworld_b200 import dist as D

    rank, world, local = D.rank_info()
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product path has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    D.init('nccl', dev)

    import gridworld_b200 as gw
    n = args.envs
    env = gw.make_vec(n, device=dev, action_space='walking', select_and_place=True, max_steps=MAX_STEPS,
                      auto_reset=True, reward_mode=args.reward_mode, grid_obs=not args.inplace)
    targets = synth_targets(n, DENSITY, args.support, 1234 + 1 + rank)
    env.set_tasks(targets.to(dev))
    P = 64
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)
    actions = torch.randint(0, 18, (P, n), device=dev, dtype=torch.int32, generator=gen)
    K, W = args.steps, args.warmup

    def barrier():
        D.barrier(dev)

    for t in range(W):
        env.step(actions[t % P])
    barrier()
    sampler = ClockSampler(physical_gpu_index(local))
    sampler.start()
    launches0 = env.gpu_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()  # CUDA events on the launching (current) stream bracket the K back-to-back step launches
    for t in range(K):
        env.step(actions[(W + t) % P])
    ev1.record()
    barrier()
    clocks = sampler.stop()
    total_ms = ev0.elapsed_time(ev1)
    launches = env.gpu_launches - launches0
    total_ms_max = D.max_over_ranks(total_ms, dev)
    value = n * world * K / (total_ms_max * 1e-3)

    # ---- e2e through the host-buffer C-ABI call (pinned host actions in, reward/done out) ----
    acts_host = actions.cpu().pin_memory()
    r_host = torch.zeros(n, dtype=torch.float64).pin_memory()
    d_host = torch.zeros(n, dtype=torch.uint8).pin_memory()
    Ke = min(K, 500)
    for t in range(3):
        env.step_host(acts_host[t % P], r_host, d_host)
    barrier()
    t0 = time.perf_counter()
    for t in range(Ke):
        env.step_host(acts_host[t % P], r_host, d_host)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    e2e_value = n * world * Ke / D.max_over_ranks(e2e_s, dev)
    # the same call with the transfers staged through cudaMemcpyAsync (reserved bit 4), for comparison
    env_st = gw.make_vec(n, device=dev, action_space='walking', select_and_place=True, max_steps=MAX_STEPS,
                         auto_reset=True, reward_mode=args.reward_mode, grid_obs=not args.inplace, _sched_bits=16)
    env_st.load_state_dict(env.state_dict())
    for t in range(3):
        env_st.step_host(acts_host[t % P], r_host, d_host)
    barrier()
    t0 = time.perf_counter()
    for t in range(Ke):
        env_st.step_host(acts_host[t % P], r_host, d_host)
    torch.cuda.synchronize(dev)
    e2e_staged = n * world * Ke / D.max_over_ranks(time.perf_counter() - t0, dev)
    del env_st

    # ---- second half of BASELINE.json's metric: max-intersection evals/sec (stateless sweep, config 4) ----
    mi_evals = None
    if not args.no_mi:
        npairs = 1_000_000
        Tp = synth_targets(npairs, DENSITY, 'compact', 1237).to(dev)
        Gp = synth_targets(npairs, DENSITY, 'compact', 1238).to(dev)
        Gp[:npairs // 2] = torch.rot90(Tp[:npairs // 2], k=1, dims=(2, 3))  # half are rotated copies
        gw.max_intersection(Gp, Tp)
        barrier()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record()
        for _ in range(5):
            gw.max_intersection(Gp, Tp)
        m1.record()
        barrier()
        mi_evals = world * npairs * 5 / (D.max_over_ranks(m0.elapsed_time(m1), dev) * 1e-3)
        del Tp, Gp

    if rank == 0:
        peak, peak_src = hbm_peak()
        bytes_per_env = A_WALK_BYTES if not args.inplace else 189
        # A step is two launches (physics k_step<0,1>, then observe k_step<0,2>, which moves ~all the bytes).
        # They are timed together: dividing the step's algorithmic bytes by the whole step duration charges the
        # physics launch to the observe kernel, i.e. the fraction below is a lower bound for the dominant kernel.
        kernel_ms = total_ms / K
        achieved = bytes_per_env * n / (kernel_ms * 1e-3) / 1e9
        line = {
            'metric': 'env_steps_per_sec', 'value': value, 'unit': 'env-steps/s', 'n_gpus': world,
            'steps': K, 'warmup': W, 'ms_per_step': total_ms_max / K, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': WORKLOAD if not args.inplace else WORKLOAD.replace(
                           'int32 grid obs materialised', 'in-place int8 state (no int32 grid copy)'),
                       'envs_per_gpu': n, 'support': args.support, 'reward_mode': args.reward_mode,
                       'parallelism': f'envs sharded x{world}, no collective',
                       'l2': f'per-step HBM traffic ~{bytes_per_env * n / 1e6:.0f} MB vs 126 MB L2 '
                             '(inputs larger than L2, no flush needed)',
                       'actions': f'{P} pre-generated device batches cycled'},
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': 'env-steps/s', 'h2d_bytes_per_step': 4 * n,
                    'd2h_bytes_per_step': 9 * n, 'steps': Ke,
                    'api': 'IGLUVecEnv.step_host -> iglu_step_walking_host (pinned host buffers, read / written by '
                           'the step kernels over PCIe)', 'staged_value': e2e_staged},
            'gpu_launches': int(launches),
            'max_intersection': {'evals_per_sec': mi_evals, 'pairs_per_gpu': 1_000_000,
                                 'inputs': 'config 4: compact targets rho=0.20; half of the grids independent, half '
                                           'a rotated copy of the target; int8 [n,9,11,11] resident in HBM'},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'traffic': recorded_traffic(),
                         'kernel': 'iglu::k_step<0,2> (observe) timed together with iglu::k_step<0,1> (physics)',
                         'kernel_ms': kernel_ms, 'launches_per_step': launches / K,
                         'algorithmic_bytes_per_launch': bytes_per_env * n, 'peak_source': peak_src},
        }
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            rate, _ = cpu_sample(2048, 20, threads)
            steps_cpu = max(20, int(rate * 8.0 / 8192))  # ~8 s of wall time on all cores
            v, dt = cpu_sample(8192, steps_cpu, threads, warmup=5)
            line['cpu_baseline'] = {'value': v, 'unit': 'env-steps/s', 'cores': threads, 'kind': 'port',
                                    'sample': f'8192 envs x {steps_cpu} steps of the workload, {dt:.1f} s wall, '
                                              'oracle/iglu_oracle.c (C port of the reference step), reward on change'}
        print(json.dumps(line), flush=True)
    D.shutdown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2000)
    ap.add_argument('--warmup', type=int, default=50)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--envs', type=int, default=ENVS_PER_GPU, help='envs per GPU')
    ap.add_argument('--support', default='compact', choices=['compact', 'full'])
    ap.add_argument('--reward-mode', default='every_step', choices=['every_step', 'on_change'])
    ap.add_argument('--inplace', action='store_true', help='do not materialise the int32 grid obs')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-mi', action='store_true', help='skip the max-intersection evals/sec leg')
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == '__main__':
    main()
