#!/usr/bin/env python
"""bench.py -- env-steps/sec of the IGLU Gridworld step path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] ...                # the CPU arm (the reference itself, all host cores)
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # one rank per GPU, envs sharded, no collective

Headline workload (config.workload): BASELINE.json configs[1] -- walking Discrete(18), 65,536 envs per GPU,
random synthetic target structures (SURVEY 8(d) generator, compact support, density 0.20: ~25 blocks and
~196 admissible (rotation, shift) pairs per task, close to the real goals.pkl mix), uniform random actions,
full maximal_intersection evaluated on EVERY step for every env whose grid is non-empty, int32 obs['grid']
materialised every step, auto-reset on done.

What is timed. Whatever --warmup says, the envs are first brought, untimed, into a DESYNCHRONISED STEADY STATE:
env e is reset at step phase[e] (uniform over one 250-step episode) of a pre-roll, so that afterwards 1/250 of the
envs reset on every step and the block-count distribution is the stationary one of the workload. (Timing the
first steps after a common reset measures empty grids: there the reward evaluation is skipped for ~98 % of the
envs.) `config.evaluated_frac` is the fraction of env-steps of the timed window in which maximal_intersection
really ran, counted on the device. The `dense` object is the same step on worlds that all hold a half-built
target (evaluated_frac = 1: what a policy that builds sees).

value  = env-steps/s with the action tensors already resident in HBM (CUDA events, max over ranks).
e2e    = the same metric through the host-buffer C-ABI call iglu_step_walking_host_async + iglu_wait: every
         step takes the int32 actions from pinned host memory and delivers reward (f64) + done (u8) into pinned
         host memory; the host needs a step's results before it issues that env group's next actions, and keeps
         two env groups in flight (EnvPool-style) so that one group's host round trip hides under the other's
         kernels. `e2e.sync_value` is the single-group, synchronise-every-step form.
roofline: HBM. achieved = A_walk * envs / mean step duration, A_walk = 5,634 algorithmic bytes per env-step
         (4,413 B of observations/results written, 1,089 B grid + 4 B action read, 64 B state read
         + 64 B written), peak = MEASURED_PEAKS.json hbm_gbs.
max_intersection: evals/s of the stateless sweep (config 4) on four workloads with an integer-pipe roofline
         object each: achieved = evals/s x P_eval (SURVEY 8(d) popcount-ANDs per evaluation, computed from the
         targets), peak = POPC thread-ops/s measured in this run by csrc/iglu_microbench.cu.
config3 / config5: BASELINE.json configs[2] (flying + target_in_obs, 262,144 envs) and configs[4] (1,048,576
         envs per GPU, in-place observations consumed by a torch policy on the device).
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
A_FLY_BYTES = 11103          # flying + target_in_obs as the reference does it (target re-copied per step)
A_FLY_OURS = 5634 + 24       # what this implementation moves: obs['target_grid'] is written once by set_tasks
A_INPLACE_BYTES = 189
ENVS_PER_GPU = 65536         # BASELINE.json configs[1]
DENSITY = 0.20
MAX_STEPS = 250
FALLBACK_HBM_GBS = 6650.0    # /opt/skills/guides/B200_PROFILING.md fallback


def synth_targets(n, rho, support, seed, device='cpu'):
    """target = (u < rho) ? 1 + floor(6 v) : 0, u, v ~ U[0,1) from a seeded torch.Generator
    (CPU generator for env tasks and parity runs; device generator for the 1 M-pair throughput sweeps)."""
    import torch
    g = torch.Generator(device).manual_seed(seed)
    u = torch.rand(n, 9, 11, 11, generator=g, device=device)
    v = torch.rand(n, 9, 11, 11, generator=g, device=device)
    t = torch.where(u < rho, 1 + torch.floor(6 * v), torch.zeros((), device=device)).to(torch.int8)
    if support == 'compact':
        m = torch.zeros(9, 11, 11, dtype=torch.bool, device=device)
        m[:5, 3:8, 3:8] = True
        t = torch.where(m, t, torch.zeros((), dtype=torch.int8, device=device))
    return t


def p_eval_mean(T):
    """SURVEY 8(d) algorithmic work of one evaluation, mean over the targets T [m, 9, 11, 11]:
    P_eval = sum_k S_k * NR_k with S_k the admissible shifts of rotation k (a rectangle given by the bounding
    box) and NR_k the non-empty (class, y, x_k) rows of the rotated target."""
    import torch
    T = T.to(torch.int16)
    occ = T != 0
    ax = occ.any(dim=1).any(dim=2)       # [m, 11] x occupied
    az = occ.any(dim=1).any(dim=1)       # [m, 11] z occupied
    idx = torch.arange(11, device=T.device)

    def span(a):
        lo = torch.where(a, idx, 11).min(1).values
        hi = torch.where(a, idx, -1).max(1).values
        return lo, hi
    xlo, xhi = span(ax)
    zlo, zhi = span(az)
    empty = ~ax.any(1)
    S = (xlo - (xhi - 10) + 1) * (zlo - (zhi - 10) + 1)
    S = torch.where(empty, torch.full_like(S, 441), S)
    nr_x = sum(((T == c).any(dim=3)).flatten(1).sum(1) for c in range(1, 7))   # rows along x: rotations 0, 2
    nr_z = sum(((T == c).any(dim=2)).flatten(1).sum(1) for c in range(1, 7))   # rows along z: rotations 1, 3
    return float((S * (2 * nr_x + 2 * nr_z)).double().mean().item())


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

    def sample(self):
        nv = self.nv
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

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            self.sample()
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


def recorded(name, key):
    """a figure taken from a committed ncu capture (profiles/<name>), or None."""
    p = os.path.join(ROOT, 'profiles', name)
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(key)
        except Exception:  # noqa: BLE001
            return None
    return None


WORKLOAD = (f'configs[1]: walking Discrete(18), {ENVS_PER_GPU} envs/GPU, synthetic compact targets rho={DENSITY} '
            f'(~25 blocks, ~196 admissible rot x shift), uniform random actions, maximal_intersection every step, '
            f'int32 grid obs materialised, auto-reset, max_steps={MAX_STEPS}, desynchronised steady state '
            f'(env e reset at a uniform phase of a {MAX_STEPS}-step pre-roll)')

PREROLL_EXTRA = 30   # steps after the last staggered reset of the pre-roll


def phases(n, seed):
    return np.random.RandomState(seed).randint(0, MAX_STEPS, n)


# --------------------------------------------------------------------------------------- CPU arms

def cpu_port_sample(n_envs, steps, threads, seed=1234 + 1, warmup=0, preroll=True):
    """Times the oracle port (C restatement of the reference step) on `n_envs` envs of the workload, from the
    same desynchronised steady state as the GPU arm."""
    from oracle import oracle as O
    T = synth_targets(n_envs, DENSITY, 'compact', seed).numpy()
    b = O.OracleBatch(n_envs, 'walking', select_and_place=True, max_steps=MAX_STEPS, threads=threads)
    b.set_tasks(T)
    rng = np.random.RandomState(0)
    acts = rng.randint(0, 18, (64, n_envs)).astype(np.int32)
    t = 0
    if preroll:
        ph = phases(n_envs, 7)
        for t in range(MAX_STEPS + PREROLL_EXTRA):
            if t < MAX_STEPS:
                b.reset((ph == t).astype(np.uint8))
            b.step_walking(acts[t % 64], auto_reset=True)
        t += 1
    for k in range(warmup):
        b.step_walking(acts[(t + k) % 64], auto_reset=True)
    t += warmup
    t0 = time.perf_counter()
    for k in range(steps):
        b.step_walking(acts[(t + k) % 64], auto_reset=True)
    dt = time.perf_counter() - t0
    evaluated = float((b.state()['grid'].reshape(n_envs, -1) != 0).any(1).mean())
    return n_envs * steps / dt, dt, evaluated


REF_DIR = os.path.join(ROOT, 'baseline', '_ref')


def python_reference_available():
    return os.path.isdir(os.path.join(REF_DIR, 'gridworld'))


def _pyref_worker(args):
    """One process = one UNMODIFIED reference env (render=False, vector_state=True) from baseline/_ref, driven
    like examples/run_env.py: random Discrete(18) actions, reset on done. Returns the wall time of `steps` steps
    after a warm-up (numba JIT) and a pre-roll to this env's own phase of the episode."""
    idx, steps, warmup = args
    os.environ['IGLU_REFERENCE_ROOT'] = REF_DIR
    sys.path.insert(0, os.path.join(ROOT, 'oracle', 'refharness'))
    import load_reference
    load_reference.load()
    import gym
    from gridworld.tasks import Task
    env = gym.make('IGLUGridworld-v0', action_space='walking', render=False, vector_state=True, size_reward=False,
                   select_and_place=True, max_steps=MAX_STEPS)
    T = synth_targets(idx + 1, DENSITY, 'compact', 1234 + 1)[idx].numpy().astype(int)
    env.set_task(Task('', T, starting_grid=[], invariant=True))
    env.reset()
    rng = np.random.RandomState(100 + idx)
    pre = 200 + int(phases(idx + 1, 7)[idx]) + warmup
    for _ in range(pre):
        _, _, done, _ = env.step(int(rng.randint(18)))
        if done:
            env.reset()
    _PYREF_BARRIER.wait()   # all workers start their timed region together
    t0 = time.perf_counter()
    for _ in range(steps):
        _, _, done, _ = env.step(int(rng.randint(18)))
        if done:
            env.reset()
    return time.perf_counter() - t0


_PYREF_BARRIER = None


def _pyref_init(barrier):
    global _PYREF_BARRIER
    _PYREF_BARRIER = barrier


def python_reference_sample(steps, warmup, procs):
    """env-steps/s of `procs` reference processes (one env each) stepping `steps` steps at the same time."""
    import multiprocessing as mp
    ctx = mp.get_context('spawn')
    with ctx.Pool(procs, initializer=_pyref_init, initargs=(ctx.Barrier(procs),)) as pool:
        walls = pool.map(_pyref_worker, [(i, steps, warmup) for i in range(procs)], chunksize=1)
    return procs * steps / max(walls), max(walls)


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # C port (always available): bounded sample, same steady state as the GPU arm
    n_envs = 8192
    rate, _, _ = cpu_port_sample(2048, 20, threads, preroll=False)
    budget_s = 60.0
    total_steps = args.steps + args.warmup + MAX_STEPS + PREROLL_EXTRA
    if n_envs / rate * total_steps > budget_s:
        n_envs = max(256, int(rate * budget_s / total_steps) // 256 * 256)
    port_v, port_dt, port_eval = cpu_port_sample(n_envs, args.steps, threads, warmup=args.warmup)
    port = {'value': port_v, 'unit': 'env-steps/s', 'cores': threads, 'kind': 'port',
            'sample': f'{n_envs} envs x {args.steps} steps, oracle/iglu_oracle.c (C restatement pinned bit-exact to '
                      f'the executed reference), one pthread per core, reward gate as in the reference',
            'nonempty_grid_frac': port_eval}
    note = ('the reference is pure Python; its unmodified package (pip --target baseline/_ref, git-ignored) is timed '
            'through its own gym.make / step API with one process per host core; cpu_port is its C restatement')
    if python_reference_available():
        value, wall = python_reference_sample(args.steps, args.warmup, threads)
        base = {'value': value, 'unit': 'env-steps/s', 'cores': threads, 'kind': 'reference',
                'sample': f'{threads} processes x 1 env x {args.steps} steps of the workload (unmodified Python '
                          f'reference from baseline/_ref, render=False, vector_state=True)'}
        ms = 1e3 * wall / args.steps
    else:
        value, base, ms = port_v, dict(port), 1e3 * port_dt / args.steps
        note = 'baseline/_ref is absent on this box: the C restatement of the reference (oracle port) is timed instead'
    line = {
        'impl': 'reference', 'metric': 'env_steps_per_sec', 'value': value, 'unit': 'env-steps/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'note': note},
        'cpu_baseline': base, 'cpu_port': port,
        'e2e': {'value': value, 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------- GPU arm

def preroll(env, actions, n, seed):
    """Desynchronise: env e is reset at pre-roll step phase[e]; afterwards 1/MAX_STEPS of the envs reset per step."""
    import torch
    ph = torch.as_tensor(phases(n, seed), device=env.device)
    P = actions.shape[0]
    for t in range(MAX_STEPS + PREROLL_EXTRA):
        if t < MAX_STEPS:
            env.reset((ph == t).to(torch.uint8))
        env.step(actions[t % P])
    return MAX_STEPS + PREROLL_EXTRA


def timed_steps(env, step_fn, K, W, dev, barrier, mid=None):
    """W untimed + K timed calls of step_fn(t); CUDA events on the launching stream. Returns (ms total,
    evaluations counted on the device during the timed window)."""
    import torch
    for t in range(W):
        step_fn(t)
    barrier()
    env.reset_evaluation_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(K):
        step_fn(W + t)
    e1.record()
    if mid is not None:
        mid()   # the GPU is still working through the enqueued steps here
    barrier()
    return e0.elapsed_time(e1), env.evaluations


def shard_state(sd, lo, hi):
    """Envs [lo, hi) of an IGLUVecEnv.state_dict() (SoA arrays are [k][n], per-env rows [n][...])."""
    out = {}
    for k, v in sd.items():
        if k in ('pose', 'init_pose', 'istate', 'tmeta'):
            out[k] = v[:, lo:hi].contiguous()
        elif k == 'was_reset':
            out[k] = v
        else:
            out[k] = v[lo:hi].contiguous()
    return out


def int_peaks():
    """thread-ops/s of the POPC(+IADD), LOP3, IADD3, IMAD chains measured by csrc/iglu_microbench.cu now."""
    import ctypes as C
    import torch
    from gridworld_b200 import build as B
    lib = C.CDLL(B.build_microbench())
    out = (C.c_double * 4)()
    lib.iglu_mb_int_peak.argtypes = [C.POINTER(C.c_double), C.c_void_p]
    torch.cuda.synchronize()
    if lib.iglu_mb_int_peak(out, None) != 0:
        return None
    return dict(zip(('popc', 'lop3', 'iadd3', 'imad'), [float(v) for v in out]))


def reward_sweep(gw, D, dev, world, barrier, peaks, name, T, G, reps=5):
    """evals/s of the stateless sweep on the pairs (G, T) + its integer-pipe roofline object."""
    import torch
    n = T.shape[0]
    gw.max_intersection(G, T)
    barrier()
    m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    m0.record()
    for _ in range(reps):
        gw.max_intersection(G, T)
    m1.record()
    barrier()
    ms = D.max_over_ranks(m0.elapsed_time(m1), dev) / reps
    evals = world * n / (ms * 1e-3)
    pe = p_eval_mean(T[:20000])
    out = {'workload': name, 'pairs_per_gpu': n, 'ms': ms, 'evals_per_sec': evals, 'p_eval_mean': pe,
           'hbm_gbs': evals / world * (2 * 1089 + 4) / 1e9}
    if peaks:
        ach = evals / world * pe
        out['roofline'] = {'bound': 'int_pipe', 'achieved': ach, 'unit': 'popcount-ANDs/s (evals/s x P_eval)',
                           'popc_peak': peaks['popc'], 'frac': ach / peaks['popc'],
                           # ncu sm__pipe_alu_cycles_active of the committed capture of this workload, where there is one
                           'alu_pipe_pct': (recorded('max_intersection_ncu.json', name) or {}).get('alu_pipe_pct')}
    return out


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
    K, W = args.steps, args.warmup
    P = 64
    extra = set(args.legs.split(',')) if args.legs else set()

    def barrier():
        D.barrier(dev)

    def make(nn, **kw):
        base = dict(device=dev, action_space='walking', select_and_place=True, max_steps=MAX_STEPS, auto_reset=True,
                    reward_mode=args.reward_mode, grid_obs=not args.inplace)
        base.update(kw)
        return gw.make_vec(nn, **base)

    env = make(n)
    targets = synth_targets(n, DENSITY, args.support, 1234 + 1 + rank).to(dev)
    env.set_tasks(targets)
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)
    actions = torch.randint(0, 18, (P, n), device=dev, dtype=torch.int32, generator=gen)
    t_off = preroll(env, actions, n, 7 + rank)
    steady = env.state_dict()

    # ---- headline: K timed steps from the desynchronised steady state ----
    sampler = ClockSampler(physical_gpu_index(local))
    sampler.start()
    launches0 = env.gpu_launches
    total_ms, evals = timed_steps(env, lambda t: env.step(actions[(t_off + t) % P]), K, W, dev, barrier,
                                  mid=sampler.sample if sampler.ok else None)
    launches = (env.gpu_launches - launches0) * K // (K + W)
    clocks = sampler.stop()
    total_ms_max = D.max_over_ranks(total_ms, dev)
    value = n * world * K / (total_ms_max * 1e-3)
    evaluated_frac = D.sum_over_ranks(evals, dev) / (n * world * K)

    # ---- e2e: host-buffer C-ABI call, two env groups in flight ----
    acts_host = actions.cpu().pin_memory()
    Ke = min(K, 500)
    h = n // 2
    groups = []
    for lo, hi in ((0, h), (h, n)):
        e = make(hi - lo)
        e.load_state_dict(shard_state(steady, lo, hi))
        groups.append((e, torch.cuda.Stream(dev), acts_host[:, lo:hi].contiguous().pin_memory(),
                       torch.zeros(hi - lo, dtype=torch.float64).pin_memory(),
                       torch.zeros(hi - lo, dtype=torch.uint8).pin_memory()))
    ret = np.zeros(1)

    def issue(g, t):
        e, s, a, r, d = g
        with torch.cuda.stream(s):
            e.step_host(a[t % P], r, d, wait=False)

    def consume(g):
        g[0].wait()
        ret[0] += float(g[3][0])   # the host touches the step's results before it issues the group's next step

    def e2e_loop(steps):
        for g in groups:
            issue(g, 0)
        for t in range(1, steps):
            for g in groups:
                consume(g)
                issue(g, t)
        for g in groups:
            consume(g)
    e2e_loop(max(3, W))
    barrier()
    t0 = time.perf_counter()
    e2e_loop(Ke)
    torch.cuda.synchronize(dev)
    e2e_value = n * world * Ke / D.max_over_ranks(time.perf_counter() - t0, dev)
    e2e_launches = sum(g[0].gpu_launches for g in groups)
    del groups
    # single group, synchronise every step (round 1's e2e definition), for comparison
    env_s = make(n)
    env_s.load_state_dict(steady)
    r_host = torch.zeros(n, dtype=torch.float64).pin_memory()
    d_host = torch.zeros(n, dtype=torch.uint8).pin_memory()
    for t in range(3):
        env_s.step_host(acts_host[t % P], r_host, d_host)
    barrier()
    t0 = time.perf_counter()
    for t in range(Ke):
        env_s.step_host(acts_host[t % P], r_host, d_host)
    torch.cuda.synchronize(dev)
    e2e_sync = n * world * Ke / D.max_over_ranks(time.perf_counter() - t0, dev)
    # the same single-group loop with the step call captured once per action buffer into a CUDA graph (the entry point
    # only enqueues work on the caller's stream, so a user can do this): one graph launch per step instead of three
    # kernel launches + the cross-stream events
    e2e_graph = None
    try:
        cap = torch.cuda.Stream(dev)
        cap.wait_stream(torch.cuda.current_stream(dev))
        graphs = []
        with torch.cuda.stream(cap):
            env_s.step_host(acts_host[0], r_host, d_host, wait=False)
            env_s.wait()
            for t in range(P):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=cap):
                    env_s.step_host(acts_host[t], r_host, d_host, wait=False)
                graphs.append(g)
            for t in range(P):   # every graph once: a graph's first launch uploads it (set-up, like the capture)
                graphs[t].replay()
                cap.synchronize()
            barrier()
            t0 = time.perf_counter()
            for t in range(Ke):
                graphs[t % P].replay()
                cap.synchronize()
                ret[0] += float(r_host[0])
            e2e_graph = n * world * Ke / D.max_over_ranks(time.perf_counter() - t0, dev)
        torch.cuda.current_stream(dev).wait_stream(cap)
        del graphs
    except Exception as exc:   # reported, not fatal: the other two forms stand
        e2e_graph = None
        sys.stderr.write(f'bench.py: graph-replayed e2e loop unavailable: {exc}\n')
    del env_s
    # both at once: two half-size groups in flight, each group's step call replayed from CUDA graphs on its own stream
    e2e_pgraph = None
    try:
        pgroups = []
        for lo, hi in ((0, h), (h, n)):
            e = make(hi - lo)
            e.load_state_dict(shard_state(steady, lo, hi))
            s = torch.cuda.Stream(dev)
            a = acts_host[:, lo:hi].contiguous().pin_memory()
            r = torch.zeros(hi - lo, dtype=torch.float64).pin_memory()
            d = torch.zeros(hi - lo, dtype=torch.uint8).pin_memory()
            s.wait_stream(torch.cuda.current_stream(dev))
            gl = []
            with torch.cuda.stream(s):
                e.step_host(a[0], r, d, wait=False)
                e.wait()
                for t in range(P):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=s):
                        e.step_host(a[t], r, d, wait=False)
                    gl.append(g)
            pgroups.append((e, s, gl, r, d, a))   # the graphs hold raw pointers into a, r, d: keep them alive
    except Exception as exc:   # reported, not fatal
        pgroups = None
        sys.stderr.write(f'bench.py: graph-replayed pipelined e2e loop unavailable: {exc}\n')
    if D.sum_over_ranks(0.0 if pgroups is None else 1.0, dev) == world:   # every rank captured its graphs

        def pg_issue(g, t):
            with torch.cuda.stream(g[1]):
                g[2][t % P].replay()

        def pg_consume(g):
            g[1].synchronize()
            ret[0] += float(g[3][0])

        def pg_loop(steps):
            for g in pgroups:
                pg_issue(g, 0)
            for t in range(1, steps):
                for g in pgroups:
                    pg_consume(g)
                    pg_issue(g, t)
            for g in pgroups:
                pg_consume(g)
        pg_loop(max(P, W))   # every graph once: a graph's first launch uploads it (set-up, like the capture)
        barrier()
        t0 = time.perf_counter()
        pg_loop(Ke)
        torch.cuda.synchronize(dev)
        e2e_pgraph = n * world * Ke / D.max_over_ranks(time.perf_counter() - t0, dev)
        for g in pgroups:
            torch.cuda.current_stream(dev).wait_stream(g[1])
    del pgroups

    # ---- dense: every world holds a half-built target, no episode limit -> every env is evaluated every step ----
    Kd = min(K, 200)
    env_d = make(n, max_steps=1 << 30)
    env_d.set_tasks(targets)
    keep = torch.rand(targets.shape, device=dev, generator=gen) < 0.5
    built = torch.where(keep, targets, torch.zeros((), dtype=torch.int8, device=dev))
    built[:, :3, 5, 5] = 0   # leave the spawn column free
    env_d.set_grids(built)
    for t in range(20):
        env_d.step(actions[t % P])
    dense_ms, dense_evals = timed_steps(env_d, lambda t: env_d.step(actions[t % P]), Kd, max(3, W), dev, barrier)
    dense_ms = D.max_over_ranks(dense_ms, dev) / Kd
    dense_frac = D.sum_over_ranks(dense_evals, dev) / (n * world * Kd)
    del env_d, built, keep

    peak, peak_src = hbm_peak()
    line = None
    if rank == 0:
        bytes_per_env = A_WALK_BYTES if not args.inplace else A_INPLACE_BYTES
        # A step is three launches on one stream: k_widen (moves ~all the bytes), the physics launch k_step<0,1> as its
        # programmatic dependent (it runs next to it), then k_commit. They are timed together: dividing the step's algorithmic bytes by the whole step
        # duration charges everything to the dominant kernel, i.e. the fraction below is a lower bound for it.
        kernel_ms = total_ms / K
        achieved = bytes_per_env * n / (kernel_ms * 1e-3) / 1e9
        line = {
            'metric': 'env_steps_per_sec', 'value': value, 'unit': 'env-steps/s', 'n_gpus': world,
            'steps': K, 'warmup': W, 'ms_per_step': total_ms_max / K, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': WORKLOAD if not args.inplace else WORKLOAD.replace(
                           'int32 grid obs materialised', 'in-place int8 state (no int32 grid copy)'),
                       'envs_per_gpu': n, 'support': args.support, 'reward_mode': args.reward_mode,
                       'evaluated_frac': evaluated_frac,
                       'preroll_steps': t_off,
                       'parallelism': f'envs sharded x{world}, no collective',
                       'l2': f'per-step HBM traffic ~{bytes_per_env * n / 1e6:.0f} MB vs 126 MB L2 '
                             '(inputs larger than L2, no flush needed)',
                       'actions': f'{P} pre-generated device batches cycled'},
            'clocks': clocks,
            'e2e': {'value': max(e2e_value, e2e_sync, e2e_graph or 0.0, e2e_pgraph or 0.0), 'unit': 'env-steps/s', 'h2d_bytes_per_step': 4 * n,
                    'd2h_bytes_per_step': 9 * n, 'steps': Ke,
                    'api': 'IGLUVecEnv.step_host -> iglu_step_walking_host[_async] / iglu_wait: pinned host actions in, '
                           'reward / done out, read / written by the step kernels over PCIe; the host reads a step\'s '
                           'results before it issues the next actions. value = the best of four host loops: '
                           f'pipelined_value (two groups of {h} envs in flight, each waits for its own results), '
                           'sync_value (one group, synchronise every step), graph_value (one group, the step call '
                           'captured into one CUDA graph per pinned action buffer and replayed, synchronise every step), '
                           'pipelined_graph_value (the two groups, each replaying its graphs on its own stream); every graph is '
                           'launched once before the timed loops (its first launch uploads it)',
                    'pipelined_value': e2e_value, 'sync_value': e2e_sync, 'graph_value': e2e_graph,
                    'pipelined_graph_value': e2e_pgraph,
                    'gpu_launches': int(e2e_launches)},
            'gpu_launches': int(launches),
            'dense': {'ms_per_step': dense_ms, 'value': n * world / (dense_ms * 1e-3), 'evaluated_frac': dense_frac,
                      'steps': Kd, 'roofline_frac': bytes_per_env * n / (dense_ms * 1e-3) / 1e9 / peak,
                      'worlds': 'each env starts from a random half of its target already built (IGLUVecEnv.set_grids), '
                                'no episode limit: a policy that builds keeps every grid non-empty'},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'traffic': recorded('step_kernel_traffic.json', 'dram_bytes_per_launch'),
                         'kernel': recorded('step_kernel_traffic.json', 'kernels'),
                         'kernel_ms': kernel_ms, 'launches_per_step': launches / K,
                         'algorithmic_bytes_per_launch': bytes_per_env * n, 'peak_source': peak_src},
        }

    # ---- config 3: flying + target_in_obs, 262,144 envs ----
    if 'noconfig3' not in extra:
        n3 = 262144
        K3 = min(K, 100)
        env3 = gw.make_vec(n3, device=dev, action_space='flying', target_in_obs=True, select_and_place=True,
                           max_steps=MAX_STEPS, auto_reset=True, reward_mode=args.reward_mode)
        env3.set_tasks(synth_targets(n3, DENSITY, args.support, 77 + rank).to(dev))
        P3 = 8
        fa = [{'movement': torch.rand(n3, 3, device=dev, generator=gen) * 2 - 1,
               'camera': torch.rand(n3, 2, device=dev, generator=gen) * 10 - 5,
               'inventory': torch.randint(0, 7, (n3,), device=dev, dtype=torch.int32, generator=gen),
               'placement': torch.randint(0, 3, (n3,), device=dev, dtype=torch.int32, generator=gen)} for _ in range(P3)]
        ph3 = torch.as_tensor(phases(n3, 11 + rank), device=dev)
        for t in range(MAX_STEPS + PREROLL_EXTRA):
            if t < MAX_STEPS:
                env3.reset((ph3 == t).to(torch.uint8))
            env3.step(fa[t % P3])
        ms3, ev3 = timed_steps(env3, lambda t: env3.step(fa[t % P3]), K3, max(3, W), dev, barrier)
        ms3 = D.max_over_ranks(ms3, dev) / K3
        ev3 = D.sum_over_ranks(ev3, dev) / (n3 * world * K3)
        if rank == 0:
            line['config3'] = {
                'workload': 'configs[2]: flying Dict + target_in_obs=True, 262,144 envs/GPU, continuous random actions, '
                            'every-step reward, int32 grid obs materialised, desynchronised steady state',
                'envs_per_gpu': n3, 'steps': K3, 'ms_per_step': ms3, 'value': n3 * world / (ms3 * 1e-3),
                'evaluated_frac': ev3,
                'roofline': {'bound': 'hbm', 'unit': 'GB/s', 'peak': peak,
                             'achieved_reference_bytes': A_FLY_BYTES * n3 / (ms3 * 1e-3) / 1e9,
                             'frac_reference_bytes': A_FLY_BYTES * n3 / (ms3 * 1e-3) / 1e9 / peak,
                             'achieved': A_FLY_OURS * n3 / (ms3 * 1e-3) / 1e9,
                             'frac': A_FLY_OURS * n3 / (ms3 * 1e-3) / 1e9 / peak,
                             'note': 'A_fly = 11,103 B/step counts the per-step re-copy of target_grid the reference does; '
                                     'here obs[target_grid] is persistent (written by set_tasks / the task generator), so '
                                     f'{A_FLY_OURS} B/step really move: `frac` uses those'}}
        del env3, fa

    # ---- obs['pov']: 64 x 64 frames of every env per step, render=True ----
    if 'nopov' not in extra:
        n_pov = min(16384, n)
        Kp = min(K, 50)
        envp = gw.make_vec(n_pov, device=dev, action_space='walking', select_and_place=True, max_steps=MAX_STEPS,
                           auto_reset=True, reward_mode=args.reward_mode, render=True)
        tp = synth_targets(n_pov, DENSITY, args.support, 55 + rank).to(dev)
        envp.set_tasks(tp, tp)       # worlds that show something: the target is already standing
        for t in range(5):
            envp.step(actions[t % P][:n_pov])
        msp, _ = timed_steps(envp, lambda t: envp.step(actions[t % P][:n_pov]), Kp, 3, dev, barrier)
        msr, _ = timed_steps(envp, lambda t: envp.render(), Kp, 3, dev, barrier)
        msp = D.max_over_ranks(msp, dev) / Kp
        msr = D.max_over_ranks(msr, dev) / Kp
        if rank == 0:
            line['pov'] = {'workload': f'render=True: {n_pov} envs/GPU, 64 x 64 x 3 uint8 frame per env and step from the '
                                       'CUDA voxel ray caster (iglu_render_pov), flat colours',
                           'envs_per_gpu': n_pov, 'steps': Kp, 'ms_per_step': msp, 'value': n_pov * world / (msp * 1e-3),
                           'render_only_ms': msr, 'frames_per_sec': n_pov * world / (msr * 1e-3),
                           'pixels_per_sec': n_pov * world * 4096 / (msr * 1e-3)}
        del envp, tp

    # ---- config 5: 1,048,576 envs per GPU, in-place observations, torch policy on the device ----
    if world > 1 or 'config5' in extra:
        n5 = 1048576
        K5 = min(K, 50)
        env5 = gw.make_vec(n5, device=dev, action_space='walking', select_and_place=True, max_steps=MAX_STEPS,
                           auto_reset=True, reward_mode='on_change', grid_obs=False, max_target_blocks=64)
        env5.set_tasks(synth_targets(n5, DENSITY, 'compact', 99 + rank, device=dev))
        torch.manual_seed(rank)
        w1 = torch.randint(-8, 9, (1104, 64), device=dev, dtype=torch.int8)   # K = the padded state row (pad cells are 0)
        w1s = torch.randn(5 + 6, 64, device=dev, dtype=torch.bfloat16) * 0.05
        w2 = torch.randn(64, 18, device=dev, dtype=torch.bfloat16) * 0.5

        def policy(o):   # the policy of examples/policy_loop.py: reads the int8 voxel state and the small obs in place
            h = torch._int_mm(env5.grid_rows, w1).to(torch.bfloat16) * 0.006   # int8 tensor cores on the int8 state itself
            h = h + torch.cat([o['agentPos'], o['inventory'] * 0.05], dim=1).to(torch.bfloat16) @ w1s
            logits = (torch.relu(h) @ w2).float()
            gumbel = -torch.log(-torch.log(torch.rand_like(logits).clamp_min(1e-20)))
            return (logits + gumbel).argmax(dim=1).to(torch.int32)   # a sample of Categorical(logits)
        obs5 = [env5.reset()]
        a5 = torch.randint(0, 18, (n5,), device=dev, dtype=torch.int32, generator=gen)
        ph5 = torch.as_tensor(phases(n5, 13 + rank), device=dev)
        for t in range(MAX_STEPS + PREROLL_EXTRA):
            if t < MAX_STEPS:
                env5.reset((ph5 == t).to(torch.uint8))
            obs5[0] = env5.step(a5)[0]

        def step5(t):
            obs5[0] = env5.step(policy(obs5[0]))[0]
        ms5, _ = timed_steps(env5, step5, K5, max(3, W), dev, barrier)
        ms5 = D.max_over_ranks(ms5, dev) / K5
        ms5e, _ = timed_steps(env5, lambda t: env5.step(a5), K5, 3, dev, barrier)
        ms5e = D.max_over_ranks(ms5e, dev) / K5
        if rank == 0:
            line['config5'] = {
                'workload': 'configs[4]: walking, 1,048,576 envs/GPU, reward on change (reference gate), observations '
                            'consumed IN PLACE by a torch MLP policy on the device (int8 first layer on the int8 voxel state rows + agentPos + '
                            'inventory -> Categorical over 18 actions), desynchronised steady state',
                'envs_per_gpu': n5, 'steps': K5, 'ms_per_step': ms5, 'value': n5 * world / (ms5 * 1e-3),
                'env_only_ms_per_step': ms5e, 'env_only_value': n5 * world / (ms5e * 1e-3),
                'hbm_frac_info': A_INPLACE_BYTES * n5 / (ms5e * 1e-3) / 1e9 / peak}
        del env5

    # ---- second half of BASELINE.json's metric: max-intersection evals/sec (stateless sweep, config 4) ----
    if not args.no_mi:
        peaks = int_peaks()
        npairs = 1_000_000
        sweeps = []
        for rho in (0.05, 0.20, 0.50):
            Tp = synth_targets(npairs, rho, 'compact', 1237, device=dev)
            Gp = synth_targets(npairs, rho, 'compact', 1238, device=dev)
            Gp[:npairs // 2] = torch.rot90(Tp[:npairs // 2], k=1, dims=(2, 3))  # half are rotated copies
            sweeps.append(reward_sweep(gw, D, dev, world, barrier, peaks, f'compact rho={rho:.2f}', Tp, Gp))
            del Tp, Gp
        for rho in (0.05, 0.20, 0.50):   # full support: targets span the zone, one shift per rotation (byte-compare paths)
            Tp = synth_targets(npairs, rho, 'full', 1237, device=dev)
            Gp = synth_targets(npairs, rho, 'full', 1238, device=dev)
            Gp[:npairs // 2] = torch.rot90(Tp[:npairs // 2], k=1, dims=(2, 3))
            sw = reward_sweep(gw, D, dev, world, barrier, peaks, f'full rho={rho:.2f}', Tp, Gp)
            sw['hbm_frac'] = sw['hbm_gbs'] / peak   # these are bandwidth- / byte-compare-bound, not shift-bound
            sweeps.append(sw)
            del Tp, Gp
        gpath = os.path.join(ROOT, 'tests', 'golden', 'goals.npz')
        if os.path.exists(gpath):
            goals = torch.as_tensor(np.load(gpath)['targets']).to(torch.int8)
            idx = torch.arange(npairs) % goals.shape[0]
            Tp = goals[idx].to(dev)
            keep = torch.rand(Tp.shape, device=dev, generator=gen) < 0.5
            Gp = torch.rot90(torch.where(keep, Tp, torch.zeros((), dtype=torch.int8, device=dev)), k=1,
                             dims=(2, 3)).contiguous()   # (a strided view would be re-packed inside every timed call)
            sweeps.append(reward_sweep(gw, D, dev, world, barrier, peaks,
                                       'goals.pkl structures tiled; grid = rotated random half of the target', Tp, Gp))
            del Tp, Gp
        if rank == 0:
            main_sw = sweeps[1]
            line['max_intersection'] = {
                'evals_per_sec': main_sw['evals_per_sec'], 'pairs_per_gpu': npairs,
                'inputs': 'config 4: compact targets rho=0.20 (headline); half of the grids independent, half a rotated '
                          'copy of the target; int8 [n,9,11,11] resident in HBM, generated on the device',
                'roofline': main_sw.get('roofline'), 'int_peaks': peaks, 'sweeps': sweeps}

    if rank == 0:
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            rate, _, _ = cpu_port_sample(2048, 20, threads, preroll=False)
            steps_cpu = max(20, int(rate * 8.0 / 8192))  # ~8 s of wall time on all cores after the pre-roll
            v, dt, nz = cpu_port_sample(8192, steps_cpu, threads, warmup=5)
            line['cpu_baseline'] = {'value': v, 'unit': 'env-steps/s', 'cores': threads, 'kind': 'port',
                                    'sample': f'8192 envs x {steps_cpu} steps of the workload from the same '
                                              f'desynchronised steady state, {dt:.1f} s wall, oracle/iglu_oracle.c (C '
                                              'port of the reference step), reward on change',
                                    'nonempty_grid_frac': nz}
            if python_reference_available() and 'nopyref' not in extra:
                pv, pw = python_reference_sample(2000, 0, threads)
                line['cpu_baseline']['python_reference'] = {
                    'value': pv, 'unit': 'env-steps/s', 'cores': threads, 'kind': 'reference',
                    'sample': f'{threads} processes x 1 env x 2000 steps, unmodified reference package (baseline/_ref) '
                              f'through gym.make / env.step, {pw:.1f} s wall'}
            if 'nodropin' not in extra:
                line['dropin_n1'] = dropin_n1(gw, dev)
        print(json.dumps(line), flush=True)
    D.shutdown()


def dropin_n1(gw, dev):
    """The N = 1 gym drop-in (gw.make(...).step, numpy observations on the host) against the reference's own
    per-step latency (~590 us mean on this workload)."""
    env = gw.make('IGLUGridworld-v0', action_space='walking', render=False, vector_state=True, size_reward=False,
                  select_and_place=True, max_steps=MAX_STEPS, device=dev)
    T = synth_targets(1, DENSITY, 'compact', 1234 + 1)[0].numpy().astype(int)
    env.set_task(gw.Task('', T, starting_grid=[]))
    env.reset()
    rng = np.random.RandomState(100)
    for _ in range(50):
        _, _, done, _ = env.step(int(rng.randint(18)))
        if done:
            env.reset()
    n = 1000
    t0 = time.perf_counter()
    for _ in range(n):
        _, _, done, _ = env.step(int(rng.randint(18)))
        if done:
            env.reset()
    us = 1e6 * (time.perf_counter() - t0) / n
    return {'us_per_step': us, 'steps': n, 'reference_us_per_step': 590.0,
            'note': 'gw.make(IGLUGridworld-v0).step incl. resets, numpy obs on the host: one step call + one '
                    'snapshot kernel + one device->host copy + one synchronisation per step; reference figure: SURVEY 6'}


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
    ap.add_argument('--legs', default='', help='comma list: config5 (force at 1 GPU), noconfig3, nopyref, nodropin, nopov')
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == '__main__':
    main()
