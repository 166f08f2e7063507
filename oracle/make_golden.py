"""Generate the golden vectors under tests/golden/ by EXECUTING the unmodified reference.

TEST INFRASTRUCTURE ONLY. Run in the build container (where /root/reference exists):

    python oracle/make_golden.py

It imports the Python reference through oracle/refharness (gym stub + numpy.int shim),
drives it with fixed-seed / scripted action traces, and records the reference's own
state after every step. Nothing here is an algorithm of ours: every number stored is an
output of the reference code (gridworld/env.py, core/world.py, tasks/task.py).

Files written (all small, committed):
  goals.npz          156 real target structures (skills/goals.pkl via Tasks.to_dense), their
                     Task.__init__ products and maximal_intersection known answers
  reward_pairs.npz   random (grid, target[, full_grid]) pairs -> Task.maximal_intersection /
                     argmax_intersection, incl. negative "difference grid" classes
  traces_walking.npz random + scripted Discrete(18) traces (per-step pose/dy/tis/inventory/
                     reward/done/grid)
  traces_flying.npz  flying Dict traces (float32 actions, widened to fp64 for the reference)
  traces_dict.npz    walking Dict (discretize=False) traces
  misc.npz           Task.get_intersection and Subtasks (reset / progressive step_intersection) known answers
  surface.npz        N = 1 gym surface: SizeReward, initialize_world / deinitialize_world, host set_task_generator
  tasks_dataset.pkl.bz2 + tasks_dataset_expected.npz   a parsed-task cache in IGLUDataset's format and the single-turn
                     tasks the reference derives from it
"""
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'refharness'))
import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
CELLS = 9 * 11 * 11


def dense_to_sparse_world(d):
    """dense [y+1, x+5, z+5] -> list of world-coordinate (x, y, z, colour), the format of
    Task.starting_grid (reference tasks/task.py:172-173)."""
    ys, xs, zs = np.nonzero(d)
    return [(int(x) - 5, int(y) - 1, int(z) - 5, int(d[y, x, z])) for y, x, z in zip(ys, xs, zs)]


# ----------------------------------------------------------------------------- reward

def synth_target(rng, rho, support):
    """SURVEY 8(d) synthetic targets: (u < rho) ? 1 + floor(6 v) : 0 on the full zone or the
    compact 5x5x5 support (y < 5, 3 <= x, z <= 7)."""
    u = rng.random_sample((9, 11, 11))
    v = rng.random_sample((9, 11, 11))
    t = np.where(u < rho, 1 + np.floor(6 * v), 0).astype(np.int8)
    if support == 'compact':
        m = np.zeros((9, 11, 11), bool)
        m[:5, 3:8, 3:8] = True
        t = np.where(m, t, 0).astype(np.int8)
    return t


def related_grid(rng, task, target):
    """A grid that is a rotated + admissibly shifted copy of the target with 10% of the cells
    re-randomised (SURVEY 8(d) config-4 inputs)."""
    k = rng.randint(4)
    adm = task.admissible[k] if len(task.admissible) > k else task.admissible[0]
    tg = np.asarray(task.target_grids[k] if len(task.admissible) > 1 else task.target_grids[0])
    dx, dz = adm[rng.randint(len(adm))] if len(adm) else (0, 0)
    g = np.zeros((9, 11, 11), np.int8)
    ys, xs, zs = np.nonzero(tg)
    for y, x, z in zip(ys, xs, zs):
        gx, gz = x - dx, z - dz
        if 0 <= gx < 11 and 0 <= gz < 11:
            g[y, gx, gz] = tg[y, x, z]
    noise = rng.random_sample(g.shape) < 0.1
    rnd = np.where(rng.random_sample(g.shape) < 0.3, rng.randint(1, 7, g.shape), 0)
    return np.where(noise, rnd, g).astype(np.int8)


def gen_goals(Task, Tasks):
    raw = open(os.path.join(load_reference.REFERENCE_ROOT, 'skills', 'goals.pkl'), 'rb').read()
    goals = pickle.loads(raw)  # opcode-audited in SURVEY section 4: numpy scalars only
    names = sorted(goals.keys(), key=lambda s: int(s[1:]))
    rng = np.random.RandomState(7)
    targets, n_adm, tsize, ka1, ka2, grids, mi, arg = [], [], [], [], [], [], [], []
    for nm in names:
        blocks = [tuple(int(v) for v in b) for b in goals[nm]]
        t = np.asarray(Tasks.to_dense(blocks)).astype(np.int8)
        task = Task('', t.astype(np.int32))
        targets.append(t)
        n_adm.append([len(a) for a in task.admissible])
        tsize.append(task.target_size)
        # KA1: grid = rot90(T, -1, (1, 2));  KA2: every 2nd block shifted +1 in x
        g1 = np.rot90(t, k=-1, axes=(1, 2)).copy()
        g2 = np.zeros_like(t)
        for i, (y, x, z) in enumerate(np.argwhere(t)):
            if i % 2 == 1:
                if x + 1 <= 10:
                    g2[y, x + 1, z] = t[y, x, z]
            else:
                g2[y, x, z] = t[y, x, z]
        ka1.append(task.maximal_intersection(g1.astype(np.int32)))
        ka2.append(task.maximal_intersection(g2.astype(np.int32)))
        gs = [related_grid(rng, task, t) for _ in range(2)]
        grids.append(gs)
        mi.append([task.maximal_intersection(g.astype(np.int32)) for g in gs])
        arg.append([task.argmax_intersection(g.astype(np.int32)) for g in gs])
    np.savez_compressed(os.path.join(OUT, 'goals.npz'), names=np.array(names), targets=np.array(targets),
                        n_adm=np.array(n_adm, np.int32), target_size=np.array(tsize, np.int32),
                        ka1=np.array(ka1, np.int32), ka2=np.array(ka2, np.int32),
                        grids=np.array(grids, np.int8), max_int=np.array(mi, np.int32),
                        argmax=np.array(arg, np.int32))
    print('goals.npz', len(names), 'structures; KA C1', ka1[0], ka2[0])


def gen_reward_pairs(Task):
    rng = np.random.RandomState(11)
    rows = []
    for rho in (0.05, 0.20, 0.50):
        for support in ('full', 'compact'):
            for j in range(10):
                t = synth_target(rng, rho, support)
                mode = j % 5  # 0,1 plain; 2 difference-grid classes; 3 full_grid; 4 invariant=False
                full = None
                invariant = True
                if mode == 2:  # values in -6..6 like env.py:230 target - init
                    start = synth_target(rng, rho / 2, support)
                    t = (t.astype(np.int32) - start).astype(np.int8)
                if mode == 3:  # full_grid = superset structure (tasks/task.py:63-66)
                    extra = synth_target(rng, 0.05, 'compact')
                    full = np.where(t != 0, t, extra).astype(np.int8)
                if mode == 4:
                    invariant = False
                task = Task('', t.astype(np.int32), full_grid=None if full is None else full.astype(np.int32),
                            invariant=invariant)
                for g in (synth_target(rng, rho, support), related_grid(rng, task, t)):
                    if mode == 2:
                        g = (g.astype(np.int32) - start).astype(np.int8)
                    rows.append((t, g, full if full is not None else np.zeros_like(t), full is not None,
                                 invariant, task.maximal_intersection(g.astype(np.int32)),
                                 task.argmax_intersection(g.astype(np.int32)),
                                 [len(a) for a in task.admissible] + [0] * (4 - len(task.admissible)),
                                 task.target_size))
    # degenerate cases: empty target (all 441 shifts admissible), single block, full-width bar
    for t in (np.zeros((9, 11, 11), np.int8),):
        task = Task('', t.astype(np.int32))
        g = synth_target(rng, 0.2, 'full')
        rows.append((t, g, np.zeros_like(t), False, True, task.maximal_intersection(g.astype(np.int32)),
                     task.argmax_intersection(g.astype(np.int32)), [len(a) for a in task.admissible],
                     task.target_size))
    # full_grid empty but target not: every shift admissible -> the slices really clip
    t = synth_target(rng, 0.2, 'full')
    full = np.zeros_like(t)
    task = Task('', t.astype(np.int32), full_grid=full.astype(np.int32))
    for _ in range(3):
        g = related_grid(rng, Task('', t.astype(np.int32)), t)
        g = np.roll(g, (rng.randint(-3, 4), rng.randint(-3, 4)), axis=(1, 2))
        rows.append((t, g, full, True, True, task.maximal_intersection(g.astype(np.int32)),
                     task.argmax_intersection(g.astype(np.int32)), [len(a) for a in task.admissible],
                     task.target_size))
    cols = list(zip(*rows))
    np.savez_compressed(os.path.join(OUT, 'reward_pairs.npz'), targets=np.array(cols[0], np.int8),
                        grids=np.array(cols[1], np.int8), fulls=np.array(cols[2], np.int8),
                        has_full=np.array(cols[3], bool), invariant=np.array(cols[4], bool),
                        max_int=np.array(cols[5], np.int32), argmax=np.array(cols[6], np.int32),
                        n_adm=np.array(cols[7], np.int32), target_size=np.array(cols[8], np.int32))
    print('reward_pairs.npz', len(rows), 'pairs')


# ----------------------------------------------------------------------------- traces

class Recorder:
    FIELDS = ('pos', 'rot', 'dy', 'tis', 'active', 'inventory', 'reward', 'done', 'max_int', 'step_no',
              'agentPos', 'compass', 'obs_inventory')

    def __init__(self):
        self.d = {k: [] for k in self.FIELDS}
        self.grid_steps, self.grid_vals = [], []
        self.last_grid = None

    def record(self, env, obs, reward, done, t):
        u = env.unwrapped
        a = u.agent
        d = self.d
        d['pos'].append([float(v) for v in a.position])
        d['rot'].append([float(v) for v in a.rotation])
        d['dy'].append(float(a.dy))
        d['tis'].append(int(a.time_int_steps))
        d['active'].append(int(a.active_block))
        d['inventory'].append([int(v) for v in a.inventory])
        d['reward'].append(float(reward))
        d['done'].append(bool(done))
        d['max_int'].append(int(u._synthetic_task.max_int))
        d['step_no'].append(int(u.step_no))
        d['agentPos'].append(np.asarray(obs['agentPos'], np.float32))
        d['compass'].append(np.asarray(obs['compass'], np.float32))
        d['obs_inventory'].append(np.asarray(obs['inventory'], np.float32))
        g = u.grid.astype(np.int8).reshape(-1)
        assert np.array_equal(np.asarray(obs['grid']).reshape(-1), g)
        if self.last_grid is None or not np.array_equal(g, self.last_grid):
            self.grid_steps.append(t)  # grid snapshots stored only when the grid changed
            self.grid_vals.append(g.copy())
            self.last_grid = g

    def pack(self, prefix, out):
        d = self.d
        out[prefix + 'pos'] = np.array(d['pos'], np.float64)
        out[prefix + 'rot'] = np.array(d['rot'], np.float64)
        out[prefix + 'dy'] = np.array(d['dy'], np.float64)
        out[prefix + 'tis'] = np.array(d['tis'], np.int32)
        out[prefix + 'active'] = np.array(d['active'], np.int32)
        out[prefix + 'inventory'] = np.array(d['inventory'], np.int32)
        out[prefix + 'reward'] = np.array(d['reward'], np.float64)
        out[prefix + 'done'] = np.array(d['done'], bool)
        out[prefix + 'max_int'] = np.array(d['max_int'], np.int32)
        out[prefix + 'step_no'] = np.array(d['step_no'], np.int32)
        out[prefix + 'agentPos'] = np.array(d['agentPos'], np.float32)
        out[prefix + 'compass'] = np.array(d['compass'], np.float32)
        out[prefix + 'obs_inventory'] = np.array(d['obs_inventory'], np.float32)
        out[prefix + 'grid_steps'] = np.array(self.grid_steps, np.int32)
        out[prefix + 'grid_vals'] = np.array(self.grid_vals, np.int8).reshape(-1, CELLS)


def make_env(gym, Task, action_space, target, start, pose, select_and_place=True, max_steps=250, discretize=True,
             right_scale=1, wrong_scale=0.1):
    env = gym.make('IGLUGridworld-v0', render=False, vector_state=True, size_reward=False,
                   select_and_place=select_and_place, action_space=action_space, max_steps=max_steps,
                   discretize=discretize, right_placement_scale=right_scale, wrong_placement_scale=wrong_scale)
    sparse = [] if start is None else dense_to_sparse_world(start)
    task = Task('', np.asarray(target).astype(np.int32), starting_grid=sparse)
    if pose is not None:
        env.unwrapped._task = task  # initialize_world() needs a task for its reset
        env.initialize_world(sparse, tuple(float(v) for v in pose))
    else:
        env.set_task(task)
    return env


def run_trace(gym, Task, cfg, actions, conv):
    """One fresh env per trace (dy / active_block / time_int_steps leak across resets).
    When the reference reports done, the harness calls reset() before the next step -- the
    per-env sequence the batched auto-reset reproduces."""
    env = make_env(gym, Task, cfg['action_space'], cfg['target'], cfg.get('start'), cfg.get('pose'),
                   cfg.get('select_and_place', True), cfg.get('max_steps', 250), cfg.get('discretize', True),
                   cfg.get('right_scale', 1), cfg.get('wrong_scale', 0.1))
    rec = Recorder()
    done = False
    n = len(actions[0]) if isinstance(actions, tuple) else len(actions)
    for t in range(n):
        if done:
            env.reset()
        obs, reward, done, _ = env.step(conv(actions, t))
        rec.record(env, obs, reward, done, t)
    return rec


def walking_action_mix(rng, n, kind):
    if kind == 'uniform':
        return rng.randint(0, 18, n).astype(np.int32)
    if kind == 'builder':  # look down, place/break/hotbar heavy, some jumps -> towers, falls, exhaustion
        p = np.array([1, 4, 2, 2, 2, 5, 2, 2, 2, 2, 2, 2, 2, 2, 5, 2, 4, 6], float)
        return rng.choice(18, n, p=p / p.sum()).astype(np.int32)
    if kind == 'runner':  # mostly movement + jumps: walls, zone border, step-ups
        p = np.array([1, 8, 3, 4, 4, 4, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 2, 2, 1, 1, 1, 2], float)
        return rng.choice(18, n, p=p / p.sum()).astype(np.int32)
    raise ValueError(kind)


def scripted_walking():
    """Scripted scenarios from SURVEY section 4 (states random actions never reach)."""
    T2 = np.zeros((9, 11, 11), np.int8)
    T2[0, 5, 3] = 1
    T2[1, 5, 3] = 1  # world (0,-1,-2) and (0,0,-2) blue
    out = []
    # known-answer physics trace (SURVEY table): free fall, forward, camera, place, break, jump
    a = [0, 0, 0, 1, 1, 1] + [15] * 9 + [14] * 18 + [17, 16, 5] + [0] * 13
    out.append(dict(name='ka_physics', cfg=dict(action_space='walking', target=T2, select_and_place=False), a=a))
    # high fall (time_int_steps 8 and 12)
    out.append(dict(name='high_fall', cfg=dict(action_space='walking', target=T2, pose=(0, 9.5, 0, 0, 0)),
                    a=[0] * 24))
    # wall
    S = np.zeros((9, 11, 11), np.int8)
    S[0, 5, 3] = 1
    S[1, 5, 3] = 1
    T2b = np.zeros((9, 11, 11), np.int8)
    T2b[0, 2, 2] = 4
    T2b[0, 5, 3] = 1
    out.append(dict(name='wall', cfg=dict(action_space='walking', target=T2b, start=S, pose=(0, -0.25, 0, 0, 0)),
                    a=[1] * 8 + [3] * 3 + [1] * 3 + [4] * 6 + [1] * 3))
    # target == start: synthetic target is empty -> target_size 0 -> done on every step
    out.append(dict(name='target_equals_start', cfg=dict(action_space='walking', target=T2, start=S,
                                                         pose=(0, -0.25, 0, 0, 0)), a=[1, 0, 17, 16, 3, 0]))
    # step-up
    S1 = np.zeros((9, 11, 11), np.int8)
    S1[0, 5, 3] = 1
    out.append(dict(name='step_up', cfg=dict(action_space='walking', target=T2, start=S1,
                                             pose=(0, -0.25, -0.75, 0, 0)), a=[5, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0]))
    # reward/done with a starting grid (SURVEY: break red start block, useless block, finish)
    S2 = np.zeros((9, 11, 11), np.int8)
    S2[0, 6, 6] = 3
    S2[0, 5, 3] = 1
    a = [14] * 9 + [13] * 27 + [16] + [12] * 27 + [17, 16] + [1, 1] + [14] * 4 + [17, 17, 0, 0]
    out.append(dict(name='start_grid_reward', cfg=dict(action_space='walking', target=T2, start=S2,
                                                       select_and_place=False), a=a))
    # tower building: look straight down, jump + place under feet repeatedly, then walk off
    # jump and place under the feet at the apex (only then the overlap test of world.py:324-325
    # lets the block through), land on it, then try again from the block (rejected: too low)
    a = [14] * 18 + [5, 0, 0, 0, 0, 17] + [0] * 12 + [5, 0, 0, 0, 0, 17] + [0] * 12 + [1] * 5 + [0] * 10
    T3 = np.zeros((9, 11, 11), np.int8)
    T3[0:5, 5, 5] = 1
    out.append(dict(name='jump_place', cfg=dict(action_space='walking', target=T3, select_and_place=False), a=a))
    # walk off a 6-block tower: fall with time_int_steps 4 and 8, land, keep walking into the border
    S4 = np.zeros((9, 11, 11), np.int8)
    S4[0:6, 5, 5] = 5
    out.append(dict(name='tower_fall', cfg=dict(action_space='walking', target=T3, start=S4,
                                                pose=(0, 5.75, 0, 0, 0)), a=[0, 0] + [1] * 6 + [0] * 25 + [16, 14, 14, 16]))
    # staircase: place a block ahead, jump on it, place the next one higher, ... (random builder
    # traces seldom get off the floor)
    a = [14] * 7 + [17] + [5, 1, 1, 1, 1, 1, 0, 0, 0] + [17] + [5, 1, 1, 1, 1, 1, 0, 0, 0] + [17] + [5, 1, 1, 1, 1, 1, 0, 0, 0] \
        + [13] * 18 + [1] * 6 + [0] * 20
    out.append(dict(name='staircase', cfg=dict(action_space='walking', target=T3, select_and_place=False), a=a))
    # leave the padded zone and come back; yaw wrap both ways; pitch clamp
    a = [1] * 40 + [13] * 80 + [2] * 10 + [12] * 160 + [3] * 40 + [15] * 40 + [14] * 50 + [4] * 60
    out.append(dict(name='zone_border_yaw_pitch', cfg=dict(action_space='walking', target=T2, max_steps=1000),
                    a=a))
    # inventory exhaustion with select_and_place hotbar: place 25 blue blocks in different spots
    a = [14] * 12
    for i in range(24):
        a += [6, 2]
    a += [7, 2, 8, 2, 16, 16]
    out.append(dict(name='inventory_exhaustion', cfg=dict(action_space='walking', target=T2, max_steps=1000),
                    a=a))
    # ceiling bump: blocks overhead
    S3 = np.zeros((9, 11, 11), np.int8)
    S3[3, 5, 5] = 2
    S3[3, 5, 4] = 2
    out.append(dict(name='ceiling', cfg=dict(action_space='walking', target=T2, start=S3), a=[0, 0, 0, 5] + [0] * 12
                    + [1, 5] + [0] * 12))
    # non multiple-of-5 initial rotation (generic trig path) and odd offsets
    out.append(dict(name='odd_rotation', cfg=dict(action_space='walking', target=T2,
                                                  pose=(0.3, -0.25, 0.1, 33.3, -41.7)),
                    a=[1, 1, 3, 17, 2, 4, 16, 13, 1, 17, 14, 14, 17, 5, 1, 1, 1]))
    # building-zone border: place / break attempts on the floor outside the 11 x 11 zone (rejected, world.py:318),
    # on the border column, breaking the WHITE / GREY floor (not removable), then the same inside
    a = [4] * 24 + [14] * 18 + [17, 16, 17] + [3] * 4 + [17, 16] + [3] * 4 + [17, 5, 17, 16, 16] + [15] * 6 + [17, 16] \
        + [2] * 30 + [17, 16, 17] + [1] * 60 + [17, 16]
    out.append(dict(name='border_place', cfg=dict(action_space='walking', target=T2, select_and_place=False), a=a))
    # spawned outside the 37 x 37 floor: nothing below, the agent falls at terminal velocity for the whole episode
    # (y ~ -600: ray-cast keys far outside any table; every action kind keeps being applied), then the auto-reset
    a = [17, 1, 16, 13, 5, 6, 14, 3, 0, 12, 9, 4, 15, 2] * 20
    out.append(dict(name='void_fall', cfg=dict(action_space='walking', target=T2, pose=(25.0, 1.0, -22.0, 45.0, -30.0)),
                    a=a[:270]))
    return out


def gen_walking(gym, Task):
    out = {}
    names = []
    rng = np.random.RandomState(0)
    idx = 0

    def add(name, cfg, a):
        nonlocal idx
        a = np.asarray(a, np.int32)
        rec = run_trace(gym, Task, cfg, a, lambda acts, t: int(acts[t]))
        p = f't{idx}_'
        out[p + 'actions'] = a
        out[p + 'target'] = np.asarray(cfg['target'], np.int8)
        out[p + 'start'] = np.asarray(cfg['start'] if cfg.get('start') is not None else np.zeros((9, 11, 11)), np.int8)
        out[p + 'pose'] = np.asarray(cfg.get('pose') or (0, 0, 0, 0, 0), np.float64)
        out[p + 'cfg'] = np.array([int(cfg.get('select_and_place', True)), int(cfg.get('max_steps', 250))], np.int32)
        out[p + 'scales'] = np.array([cfg.get('right_scale', 1), cfg.get('wrong_scale', 0.1)], np.float64)
        rec.pack(p, out)
        names.append(name)
        idx += 1
        print('  walking', name, len(a), 'steps; grid changes', len(rec.grid_steps), 'dones', int(np.sum(rec.d['done'])))

    for s in scripted_walking():
        add(s['name'], s['cfg'], s['a'])
    for kind, n in (('uniform', 1500), ('builder', 1500), ('runner', 1200)):
        for rep in range(2):
            rho = (0.05, 0.2)[rep]
            target = synth_target(rng, rho, 'compact')
            start = synth_target(rng, 0.03, 'compact') if rep == 1 else None
            if start is not None:
                start[:, 5, 5] = 0  # keep the spawn column free
                start[2:] = 0
            cfg = dict(action_space='walking', target=target, start=start, select_and_place=(kind != 'runner'),
                       max_steps=250 if rep == 0 else 400, right_scale=1 if rep == 0 else 2.0,
                       wrong_scale=0.1 if rep == 0 else 0.25)
            add(f'{kind}_{rep}', cfg, walking_action_mix(rng, n, kind))
    out['names'] = np.array(names)
    np.savez_compressed(os.path.join(OUT, 'traces_walking.npz'), **out)
    print('traces_walking.npz', idx, 'traces')


def gen_flying(gym, Task):
    out = {}
    names = []
    rng = np.random.RandomState(1)
    idx = 0

    def conv(acts, t):
        mv, cam, inv, pl = acts
        # float32 -> Python float: numpy-2 would otherwise keep float32 arithmetic
        return {'movement': [float(v) for v in mv[t]], 'camera': [float(v) for v in cam[t]],
                'inventory': int(inv[t]), 'placement': int(pl[t])}

    def add(name, cfg, n, style):
        nonlocal idx
        mv = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
        cam = rng.uniform(-5, 5, (n, 2)).astype(np.float32)
        inv = rng.randint(0, 7, n).astype(np.int32)
        pl = rng.randint(0, 3, n).astype(np.int32)
        if style == 'sparse_moves':  # exact zeros in movement components exercise the strafe branches
            mv[rng.random_sample((n, 3)) < 0.5] = 0
            cam[rng.random_sample((n, 2)) < 0.5] = 0
        if style == 'look_down':
            cam[:, 1] = -np.abs(cam[:, 1])
            mv[:, 2] *= 0.2
        acts = (mv, cam, inv, pl)
        rec = run_trace(gym, Task, cfg, acts, conv)
        p = f't{idx}_'
        out[p + 'movement'], out[p + 'camera'], out[p + 'inv'], out[p + 'placement'] = mv, cam, inv, pl
        out[p + 'target'] = np.asarray(cfg['target'], np.int8)
        out[p + 'start'] = np.asarray(cfg['start'] if cfg.get('start') is not None else np.zeros((9, 11, 11)), np.int8)
        out[p + 'pose'] = np.asarray(cfg.get('pose') or (0, 0, 0, 0, 0), np.float64)
        out[p + 'cfg'] = np.array([int(cfg.get('select_and_place', True)), int(cfg.get('max_steps', 250))], np.int32)
        out[p + 'scales'] = np.array([cfg.get('right_scale', 1), cfg.get('wrong_scale', 0.1)], np.float64)
        rec.pack(p, out)
        names.append(name)
        idx += 1
        print('  flying', name, n, 'steps; grid changes', len(rec.grid_steps), 'dones', int(np.sum(rec.d['done'])))

    for i, style in enumerate(('uniform', 'sparse_moves', 'look_down', 'look_down')):
        target = synth_target(rng, 0.1, 'compact')
        start = synth_target(rng, 0.05, 'compact') if i >= 2 else None
        cfg = dict(action_space='flying', target=target, start=start, select_and_place=(i != 1),
                   pose=(0, 1.5, 0, 0, -30) if i == 3 else None)
        add(f'{style}_{i}', cfg, 800, style)
    out['names'] = np.array(names)
    np.savez_compressed(os.path.join(OUT, 'traces_flying.npz'), **out)
    print('traces_flying.npz', idx, 'traces')


def gen_dict(gym, Task):
    out = {}
    names = []
    rng = np.random.RandomState(2)
    KEYS = ('forward', 'back', 'left', 'right', 'jump', 'attack', 'use', 'hotbar')

    def conv(acts, t):
        b, cam = acts
        a = {k: int(b[t, i]) for i, k in enumerate(KEYS)}
        a['camera'] = [float(v) for v in cam[t]]
        return a

    for idx, snap in enumerate((True, False)):
        n = 800
        b = (rng.random_sample((n, 8)) < 0.25).astype(np.int32)
        b[:, 7] = np.where(rng.random_sample(n) < 0.3, rng.randint(1, 7, n), 0)
        cam = rng.uniform(-5, 5, (n, 2)).astype(np.float32)
        if snap:  # camera in multiples of 5 -> LUT trig path, combined strafes give +-45/135 deg
            cam = (np.round(cam / 5) * 5).astype(np.float32)
        target = synth_target(rng, 0.1, 'compact')
        cfg = dict(action_space='walking', discretize=False, target=target, select_and_place=bool(idx))
        rec = run_trace(gym, Task, cfg, (b, cam), conv)
        p = f't{idx}_'
        out[p + 'buttons'], out[p + 'camera'] = b, cam
        out[p + 'target'] = target
        out[p + 'start'] = np.zeros((9, 11, 11), np.int8)
        out[p + 'pose'] = np.zeros(5)
        out[p + 'cfg'] = np.array([int(cfg['select_and_place']), 250], np.int32)
        out[p + 'scales'] = np.array([1, 0.1], np.float64)
        rec.pack(p, out)
        names.append(f'dict_{idx}')
        print('  dict', idx, n, 'steps; grid changes', len(rec.grid_steps), 'dones', int(np.sum(rec.d['done'])))
    out['names'] = np.array(names)
    np.savez_compressed(os.path.join(OUT, 'traces_dict.npz'), **out)


def gen_misc(Task, Tasks):
    """get_intersection / Subtasks (reference tasks/task.py:138-145, 208-308) known answers."""
    from gridworld.tasks.task import Subtasks
    rng = np.random.RandomState(21)
    out = {}
    # get_intersection on admissible and non-admissible (clipping) shifts
    T = synth_target(rng, 0.2, 'compact')
    task = Task('', T.astype(np.int32))
    G = related_grid(rng, task, T)
    q = [(dx, dz, r) for r in range(4) for dx in (-4, -1, 0, 2, 5) for dz in (-3, 0, 1, 6)]
    out['gi_target'], out['gi_grid'] = T, G
    out['gi_query'] = np.array(q, np.int32)
    out['gi_value'] = np.array([task.get_intersection(G.astype(np.int32), dx, dz, r) for dx, dz, r in q], np.int32)
    # Subtasks: three stages, progressive goals
    seq = [[(0, -1, 0, 1), (1, -1, 0, 1)],
           [(0, -1, 0, 1), (1, -1, 0, 1), (1, 0, 0, 2)],
           [(0, -1, 0, 1), (1, -1, 0, 1), (1, 0, 0, 2), (1, 1, 0, 3), (-2, -1, 3, 4)]]
    dialog = ['<A> build a base', '<A> put a green one on top', ['<A> finish it', '<B> ok']]
    np.random.seed(5)
    st = Subtasks(dialog, seq, invariant=False, progressive=True)
    rec = []
    for ep in range(6):
        cur = st.reset()
        rec.append((st.task_start, st.task_goal, np.asarray(cur.target_grid).astype(np.int8),
                    np.asarray(Tasks.to_dense(cur.starting_grid)).astype(np.int8), cur.max_int, cur.target_size,
                    cur.chat))
    out['st_start'] = np.array([r[0] for r in rec], np.int32)
    out['st_goal'] = np.array([r[1] for r in rec], np.int32)
    out['st_target'] = np.array([r[2] for r in rec])
    out['st_init'] = np.array([r[3] for r in rec])
    out['st_max_int'] = np.array([r[4] for r in rec], np.int32)
    out['st_target_size'] = np.array([r[5] for r in rec], np.int32)
    out['st_chat'] = np.array([r[6] for r in rec])
    # progressive step_intersection: feed the grids of the successive stages
    np.random.seed(7)
    st = Subtasks(dialog, seq, invariant=False, progressive=True)
    st.next = -1  # start from the empty world
    st.reset()
    steps = []
    for blocks in seq:
        g = np.asarray(Tasks.to_dense(blocks)).astype(np.int32)
        r, w, d = st.step_intersection(g)
        steps.append((r, w, int(d), st.task_goal))
    out['st_steps'] = np.array(steps, np.int32)
    np.savez_compressed(os.path.join(OUT, 'misc.npz'), **out)
    print('misc.npz', out['gi_value'][:8], out['st_start'], out['st_goal'], out['st_steps'].tolist())


def gen_generator(gym, Task, Tasks):
    """env.set_task_generator (reference env.py:166-172, 213-216): every reset first asks the generator for
    the next task. The generator is a finite pool drawn with np.random.choice, like RandomTasks(max_cache=m)
    .reset (task_set.py:120-127) -- but with starting_grid=[] / real starting grids, because a Task whose
    starting_grid is None crashes GridWorld.step. The chosen indices are recorded so that the
    batched build can be driven through the same sequence."""
    rng = np.random.RandomState(5)
    out, names = {}, []

    class PoolGen(Tasks):
        def __init__(self, tasks):
            self.tasks = tasks
            self.choices = []
            self.current = None

        def reset(self):
            j = int(np.random.choice(len(self.tasks)))
            self.choices.append(j)
            self.current = self.tasks[j]
            return self.current

    # where does "select BLUE, look down 45 degrees, place" put its block from the spawn pose? (dry run)
    SCRIPT = [6] + [14] * 9 + [17]
    env = gym.make('IGLUGridworld-v0', render=False, vector_state=True, size_reward=False, action_space='walking')
    far = np.zeros((9, 11, 11), np.int32)
    far[0, 0, 0] = 3
    env.set_task(Task('', far, starting_grid=[]))
    for a in SCRIPT:
        env.step(a)
    tiny = env.unwrapped.grid.astype(np.int8).copy()
    assert np.count_nonzero(tiny) == 1, np.argwhere(tiny)
    print('  scripted placement lands at dense', np.argwhere(tiny)[0], 'colour', int(tiny.max()))

    def pool(m, with_start):
        ts, ss = [], []
        for j in range(m):
            t = synth_target(rng, (0.03, 0.08, 0.2)[j % 3], 'compact')
            t[0, 5, 6] = 1 + j % 6  # one block right in front of the spawn point: reachable 'done' for tiny targets
            s = None
            if with_start and j % 2 == 1:
                s = synth_target(rng, 0.04, 'compact')
                s[:, 5, 5] = 0
                s[2:] = 0
            ts.append(t)
            ss.append(np.zeros((9, 11, 11), np.int8) if s is None else s)
        ts[0] = tiny.copy()  # a one-block task: completed (done) by the scripted placement
        ss[0] = np.zeros((9, 11, 11), np.int8)
        return np.array(ts, np.int8), np.array(ss, np.int8)

    idx = 0
    for name, space, n, max_steps, with_start in (('walking_pool', 'walking', 900, 30, True),
                                                  ('walking_pool_nostart', 'walking', 600, 25, False),
                                                  ('flying_pool', 'flying', 400, 40, True)):
        ts, ss = pool(7, with_start)
        tasks = [Task('', ts[j].astype(np.int32), starting_grid=dense_to_sparse_world(ss[j])) for j in range(len(ts))]
        gen = PoolGen(tasks)
        np.random.seed(100 + idx)  # the generator draws from numpy's global stream, like the reference's
        env = gym.make('IGLUGridworld-v0', render=False, vector_state=True, size_reward=False, target_in_obs=True,
                       action_space=space, max_steps=max_steps)
        env.set_task_generator(gen)
        rec = Recorder()
        if space == 'walking':
            a = walking_action_mix(rng, n, 'builder')
            conv = lambda acts, t: int(acts[t])  # noqa: E731
            acts = a
            out[f't{idx}_actions'] = a
        else:
            mv = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
            cam = rng.uniform(-5, 5, (n, 2)).astype(np.float32)
            inv = rng.randint(0, 7, n).astype(np.int32)
            pl = rng.choice([0, 0, 1, 1, 2], n).astype(np.int32)
            acts = (mv, cam, inv, pl)
            conv = lambda ac, t: {'movement': [float(v) for v in ac[0][t]], 'camera': [float(v) for v in ac[1][t]],  # noqa: E731
                                  'inventory': int(ac[2][t]), 'placement': int(ac[3][t])}
            out[f't{idx}_movement'], out[f't{idx}_camera'] = mv, cam
            out[f't{idx}_inv'], out[f't{idx}_placement'] = inv, pl
        done = False
        task_at_step = []
        ep_t = 0
        for t in range(n):
            if done:
                env.reset()
                ep_t = 0
            if name == 'walking_pool' and ep_t < len(SCRIPT):
                acts[t] = SCRIPT[ep_t]  # every episode opens with the scripted placement (completes the tiny task)
            ep_t += 1
            obs, reward, done, _ = env.step(conv(acts, t))
            j = gen.choices[-1]
            assert np.array_equal(np.asarray(obs['target_grid']), ts[j].astype(np.int32))
            task_at_step.append(j)
            rec.record(env, obs, reward, done, t)
        p = f't{idx}_'
        out[p + 'pool_targets'], out[p + 'pool_starts'] = ts, ss
        out[p + 'choices'] = np.array(gen.choices, np.int32)
        out[p + 'task_at_step'] = np.array(task_at_step, np.int32)
        out[p + 'cfg'] = np.array([1, max_steps], np.int32)
        out[p + 'space'] = np.array(space)
        rec.pack(p, out)
        names.append(name)
        print('  generator', name, n, 'steps; resets', len(gen.choices), 'dones', int(np.sum(rec.d['done'])),
              'early dones', int(np.sum(np.array(rec.d['done']) & (np.array(rec.d['step_no']) < max_steps))))
        idx += 1
    out['names'] = np.array(names)
    np.savez_compressed(os.path.join(OUT, 'traces_generator.npz'), **out)
    print('traces_generator.npz', idx, 'traces')


def gen_surface(gym, Task, Tasks):
    """The N = 1 gym surface beyond set_task / reset / step, recorded from the reference:
    the SizeReward wrapper (env.py:316-331; its reward only ever sees GridWorld.max_int / .wrong_placement as reset
    left them), initialize_world / deinitialize_world (env.py:177-193) and a host-side set_task_generator
    (env.py:168-175, 213-216) whose generator hands its tasks out in a fixed cycle."""
    rng = np.random.RandomState(77)
    out = {}
    T = np.zeros((9, 11, 11), np.int8)
    T[0, 5, 6] = 1
    T[0, 6, 6] = 3
    T[0, 4, 6] = 2
    T[1, 5, 6] = 1
    S = np.zeros((9, 11, 11), np.int8)
    S[0, 6, 6] = 3            # one target block is there from the start: max_int at reset = 1
    S[0, 2, 2] = 4            # ... and one that is not part of the target

    def state(env):
        u = env.unwrapped
        return ([float(v) for v in u.agent.position] + [float(v) for v in u.agent.rotation],
                [int(v) for v in u.agent.inventory], u.grid.astype(np.int8).reshape(-1).copy())

    # ---- A: SizeReward ----
    env = gym.make('IGLUGridworld-v0', render=False, vector_state=True, size_reward=True, action_space='walking',
                   max_steps=40)
    env.set_task(Task('size', T.astype(np.int32), starting_grid=dense_to_sparse_world(S)))
    acts = walking_action_mix(rng, 150, 'builder')
    acts[:11] = [6] + [14] * 9 + [17]   # the scripted placement of gen_generator: puts BLUE at dense (0, 5, 6)
    rew, dn, pose, inv, grids, resets = [], [], [], [], [], []
    done = False
    for t, a in enumerate(acts):
        if done:
            env.reset()
        resets.append(done)
        obs, r, done, _ = env.step(int(a))
        p, i, g = state(env)
        rew.append(float(r)); dn.append(bool(done)); pose.append(p); inv.append(i); grids.append(g)
    out['size_target'], out['size_start'], out['size_actions'] = T, S, acts
    out['size_reward'], out['size_done'] = np.array(rew, np.float64), np.array(dn, bool)
    out['size_pose'], out['size_inv'] = np.array(pose, np.float64), np.array(inv, np.int32)
    out['size_grid'], out['size_resets'] = np.array(grids, np.int8), np.array(resets, bool)
    out['size_max_int_attr'] = np.array(int(env.unwrapped.max_int), np.int32)
    print('  SizeReward rewards seen:', sorted(set(rew)), 'dones', int(np.sum(dn)))

    # ---- B: initialize_world / deinitialize_world ----
    env = gym.make('IGLUGridworld-v0', render=False, vector_state=True, size_reward=False, action_space='walking',
                   max_steps=60)
    env.set_task(Task('init', T.astype(np.int32), starting_grid=dense_to_sparse_world(S)))
    W = np.zeros((9, 11, 11), np.int8)
    W[0, 7, 7] = 5
    W[0, 5, 6] = 1            # a target block in the overwritten world
    W[1, 3, 3] = 6
    segs = []
    acts = walking_action_mix(rng, 120, 'builder')
    rew, dn, pose, inv, grids, apos, ogrid = [], [], [], [], [], [], []
    events = []               # (step index, kind): 1 = initialize_world before this step, 2 = deinitialize_world, 3 = reset
    done = False
    for t, a in enumerate(acts):
        if t == 20:
            env.initialize_world(dense_to_sparse_world(W), (1.5, 0.25, -2.0, 45.0, -10.0))
            events.append((t, 1)); done = False
        elif t == 75:
            env.deinitialize_world()
            events.append((t, 2)); done = False
        elif done:
            env.reset(); events.append((t, 3))
        obs, r, done, _ = env.step(int(a))
        p, i, g = state(env)
        rew.append(float(r)); dn.append(bool(done)); pose.append(p); inv.append(i); grids.append(g)
        apos.append(np.asarray(obs['agentPos'], np.float32)); ogrid.append(np.asarray(obs['grid'], np.int32).reshape(-1))
    out['iw_target'], out['iw_start'], out['iw_world'], out['iw_actions'] = T, S, W, acts
    out['iw_pose0'] = np.array([1.5, 0.25, -2.0, 45.0, -10.0])
    out['iw_events'] = np.array(events, np.int32)
    out['iw_reward'], out['iw_done'] = np.array(rew, np.float64), np.array(dn, bool)
    out['iw_pose'], out['iw_inv'] = np.array(pose, np.float64), np.array(inv, np.int32)
    out['iw_grid'], out['iw_agentPos'] = np.array(grids, np.int8), np.array(apos, np.float32)
    print('  initialize_world events', events, 'reward sum', float(np.sum(rew)))

    # ---- C: set_task_generator with a host-side generator that cycles through its tasks ----
    class Cycle(Tasks):
        def __init__(self, tasks):
            self.tasks, self.i, self.current = tasks, 0, None

        def reset(self):
            self.current = self.tasks[self.i % len(self.tasks)]
            self.i += 1
            return self.current

    pool_t, pool_s = [], []
    for j in range(4):
        t = synth_target(rng, 0.08, 'compact')
        t[0, 5, 6] = 1 + j
        s = synth_target(rng, 0.04, 'compact') if j % 2 else np.zeros((9, 11, 11), np.int8)
        s[:, 5, 5] = 0
        s[2:] = 0
        pool_t.append(t); pool_s.append(s)
    one = np.zeros((9, 11, 11), np.int8)
    one[0, 5, 6] = 1          # finished by the scripted placement -> an early done
    pool_t[0], pool_s[0] = one, np.zeros((9, 11, 11), np.int8)
    tasks = [Task(f'task {j}', pool_t[j].astype(np.int32), starting_grid=dense_to_sparse_world(pool_s[j]))
             for j in range(4)]
    env = gym.make('IGLUGridworld-v0', render=False, vector_state=True, size_reward=False, action_space='walking',
                   max_steps=25, target_in_obs=True)
    env.set_task_generator(Cycle(tasks))
    acts = walking_action_mix(rng, 160, 'builder')
    rew, dn, pose, inv, grids, dialog, tsum, resets = [], [], [], [], [], [], [], []
    done, ep_t = False, 0
    SCRIPT = [6] + [14] * 9 + [17]
    for t in range(len(acts)):
        if done:
            env.reset(); ep_t = 0
        resets.append(done)
        if ep_t < len(SCRIPT):
            acts[t] = SCRIPT[ep_t]
        ep_t += 1
        obs, r, done, _ = env.step(int(acts[t]))
        p, i, g = state(env)
        rew.append(float(r)); dn.append(bool(done)); pose.append(p); inv.append(i); grids.append(g)
        dialog.append(obs['dialog']); tsum.append(np.asarray(obs['target_grid'], np.int32).reshape(-1))
    out['gen_targets'], out['gen_starts'], out['gen_actions'] = np.array(pool_t), np.array(pool_s), acts
    out['gen_reward'], out['gen_done'] = np.array(rew, np.float64), np.array(dn, bool)
    out['gen_pose'], out['gen_inv'] = np.array(pose, np.float64), np.array(inv, np.int32)
    out['gen_grid'], out['gen_dialog'] = np.array(grids, np.int8), np.array(dialog)
    out['gen_target_obs'], out['gen_resets'] = np.array(tsum, np.int32), np.array(resets, bool)
    print('  host generator: dones', int(np.sum(dn)), 'early', int(np.sum(np.array(dn) & (np.arange(len(dn)) % 25 != 24))),
          'dialogs', sorted(set(dialog)))
    np.savez_compressed(os.path.join(OUT, 'surface.npz'), **out)
    print('surface.npz')


def gen_staged(gym, Task, Tasks):
    """Multi-turn structures with progressive goals (reference tasks/task.py:208-308). GridWorld.step itself only ever
    scores the single synthetic task of the episode (env.py:291); Subtasks.step_intersection -- which moves on to the
    next turn's target when the current one is complete -- is the call a training loop makes on the env's grid. The
    harness does exactly that after every env.step and records what the reference's Subtasks returned."""
    from gridworld.tasks.task import Subtasks
    A = [[(0, -1, -1, 1)],
         [(0, -1, -1, 1), (1, -1, -1, 2)],
         [(0, -1, -1, 1), (1, -1, -1, 2), (1, -1, 0, 3)]]
    B = [[(0, -1, -1, 1)]]
    dlgA = ['<A> a blue one', ['<A> a green one next to it', '<B> ok'], '<A> and a red one']
    np.random.seed(0)
    stA = Subtasks(dlgA, A)
    stB = Subtasks(['<A> just one'], B)
    rng = np.random.RandomState(41)
    S1, S2, S3 = [6] + [14] * 9 + [17], [13] * 9 + [7], [13] * 9 + [8]
    episodes = [
        (0, -1, S1 + S2 + [0, 16] + S3 + [12] * 18 + [16] + list(walking_action_mix(rng, 15, 'builder'))),
        (0, 0, [14] * 9 + S2 + S3 + list(walking_action_mix(rng, 25, 'builder'))),
        (1, -1, S1 + list(walking_action_mix(rng, 12, 'builder'))),
        (0, 1, list(walking_action_mix(rng, 45, 'builder'))),
    ]
    env = gym.make('IGLUGridworld-v0', render=False, vector_state=True, size_reward=False, action_space='walking',
                   max_steps=1000)
    rows = []
    first = True
    for ep, (which, turn, acts) in enumerate(episodes):
        st = (stA, stB)[which]
        st.next = turn
        if first or env.unwrapped._task is not st:
            env.set_task(st)          # resets: Subtasks.reset() -> create_task(turn, turn + 1)
            first = False
        else:
            env.reset()
        assert (st.task_start, st.task_goal) == (turn, turn + 1)
        for a in acts:
            obs, r, done, _ = env.step(int(a))
            g = env.unwrapped.grid.copy()
            sr, sw, sd = st.step_intersection(g)
            rows.append((ep, which, turn, int(a), int(sr), int(sw), int(bool(sd)), int(st.task_goal),
                         int(st.current.max_int), float(r), g.astype(np.int8).reshape(-1)))
    out = dict(ep=np.array([r[0] for r in rows], np.int32), structure=np.array([r[1] for r in rows], np.int32),
               turn=np.array([r[2] for r in rows], np.int32), actions=np.array([r[3] for r in rows], np.int32),
               right=np.array([r[4] for r in rows], np.int32), wrong=np.array([r[5] for r in rows], np.int32),
               done=np.array([r[6] for r in rows], np.int32), goal=np.array([r[7] for r in rows], np.int32),
               max_int=np.array([r[8] for r in rows], np.int32), env_reward=np.array([r[9] for r in rows]),
               grid=np.array([r[10] for r in rows], np.int8))
    for name, seq in (('A', A), ('B', B)):
        out['seq' + name] = np.array([np.asarray(Tasks.to_dense(list(b))).astype(np.int8) for b in seq])
    np.savez_compressed(os.path.join(OUT, 'staged.npz'), **out)
    print('staged.npz', len(rows), 'steps; goals reached', sorted(set(out['goal'].tolist())), 'dones',
          int(out['done'].sum()), 'rights', int((out['right'] != 0).sum()), 'wrongs', int((out['wrong'] != 0).sum()))


def gen_dataset(Tasks):
    """A parsed-task cache as IGLUDataset.dump_tasks_dataset writes it (data/iglu_dataset.py:204-209: bz2-compressed
    pickle of {task id: [Subtasks, ...]}) with the single-turn tasks the reference's Subtasks.__iter__ derives from
    it, for gridworld_b200.tasks.load_tasks_dataset."""
    import bz2
    from gridworld.tasks.task import Subtasks
    rng = np.random.RandomState(31)
    np.random.seed(3)
    data = {}
    for name, turns in (('c17', 4), ('c3', 1), ('c101', 3)):
        sessions = []
        for s_i in range(2 if turns > 1 else 1):
            blocks, seq, dialog = [], [], []
            for t in range(turns):
                for _ in range(rng.randint(1, 4)):
                    cell = (int(rng.randint(-5, 6)), int(rng.randint(-1, 4)), int(rng.randint(-5, 6)))
                    if all(b[:3] != cell for b in blocks):
                        blocks.append(cell + (int(rng.randint(1, 7)),))
                if t == 2 and blocks:
                    blocks.pop(0)      # a turn that also removes a block
                seq.append(list(blocks))
                dialog.append([f'<A> turn {t} of {name}', '<B> ok'] if t % 2 else f'<A> turn {t} of {name}')
            sessions.append(Subtasks(dialog, seq, invariant=False, progressive=True))
        data[name] = sessions
    with open(os.path.join(OUT, 'tasks_dataset.pkl.bz2'), 'wb') as f:
        f.write(bz2.compress(pickle.dumps(data)))
    rows = []
    for name in sorted(data):
        for k, st in enumerate(data[name]):
            for task in st:
                rows.append((name, k, np.asarray(task.target_grid).astype(np.int8),
                             np.asarray(Tasks.to_dense(task.starting_grid)).astype(np.int8),
                             np.asarray(task.full_grid).astype(np.int8), task.chat))
    np.savez_compressed(os.path.join(OUT, 'tasks_dataset_expected.npz'),
                        names=np.array([r[0] for r in rows]), session=np.array([r[1] for r in rows], np.int32),
                        targets=np.array([r[2] for r in rows]), starts=np.array([r[3] for r in rows]),
                        fulls=np.array([r[4] for r in rows]), chats=np.array([r[5] for r in rows]))
    print('tasks_dataset.pkl.bz2', len(rows), 'single-turn tasks')


def main():
    load_reference.load()
    import gym
    from gridworld.tasks import Task, Tasks
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ['goals', 'pairs', 'walking', 'flying', 'dict', 'misc', 'generator', 'surface', 'dataset', 'staged']
    if 'goals' in which:
        gen_goals(Task, Tasks)
    if 'pairs' in which:
        gen_reward_pairs(Task)
    if 'walking' in which:
        gen_walking(gym, Task)
    if 'flying' in which:
        gen_flying(gym, Task)
    if 'dict' in which:
        gen_dict(gym, Task)
    if 'misc' in which:
        gen_misc(Task, Tasks)
    if 'generator' in which:
        gen_generator(gym, Task, Tasks)
    if 'surface' in which:
        gen_surface(gym, Task, Tasks)
    if 'dataset' in which:
        gen_dataset(Tasks)
    if 'staged' in which:
        gen_staged(gym, Task, Tasks)


if __name__ == '__main__':
    main()
