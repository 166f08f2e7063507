"""GPU parity tests (run on the B200 box: `pytest -m gpu`). They call the CUDA path through the C ABI
(via gridworld_b200's ctypes binding) and compare with (a) the golden vectors recorded from the
executed reference and (b) the C oracle on the same seeded inputs.

Bars: grids / inventories / rewards / done / step counters / max_int bit-exact; pose (fp64 position,
rotation, dy) bit-exact on the Discrete(18) path (libm trig table), within 1e-6 on paths that use the
device sincos/atan2 (flying, non-multiple-of-5 angles), as BASELINE.json's north_star states.
"""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from tests import golden_io as G

pytestmark = pytest.mark.gpu

POSE_TOL = 1e-6


def _make(tr, kind, n=1, **kw):
    import gridworld_b200 as gw
    if kind == 'flying':
        space, disc = 'flying', True
    elif kind == 'dict':
        space, disc = 'walking', False
    else:
        space, disc = 'walking', True
    return gw.make_vec(n, action_space=space, discretize=disc, select_and_place=tr['select_and_place'],
                       max_steps=tr['max_steps'], right_placement_scale=tr['right_scale'],
                       wrong_placement_scale=tr['wrong_scale'], auto_reset=True, **kw)


def _replay_gpu(tr, kind, reward_mode='on_change', exact_pose=True, bits=0):
    env = _make(tr, kind, reward_mode=reward_mode, _sched_bits=bits)
    dev = env.device
    env.set_tasks(tr['target'][None], tr['start'][None], tr['pose'][None])
    T = tr['n']
    a = tr['actions']
    if kind == 'walking':
        acts = torch.as_tensor(a['actions'], device=dev).view(T, 1)
    elif kind == 'flying':
        mv = torch.as_tensor(a['movement'], device=dev).view(T, 1, 3)
        cam = torch.as_tensor(a['camera'], device=dev).view(T, 1, 2)
        inv = torch.as_tensor(a['inv'], device=dev).view(T, 1)
        pl = torch.as_tensor(a['placement'], device=dev).view(T, 1)
    else:
        bt = torch.as_tensor(a['buttons'], device=dev).view(T, 1, 8)
        cam = torch.as_tensor(a['camera'], device=dev).view(T, 1, 2)
    H = dict(pose=torch.zeros(T, 6, dtype=torch.float64, device=dev),
             ist=torch.zeros(T, 8, dtype=torch.int32, device=dev),
             reward=torch.zeros(T, dtype=torch.float64, device=dev),
             done=torch.zeros(T, dtype=torch.uint8, device=dev),
             grid=torch.zeros(T, 1089, dtype=torch.int8, device=dev),
             ogrid=torch.zeros(T, 1089, dtype=torch.int32, device=dev),
             apos=torch.zeros(T, 5, dtype=torch.float32, device=dev),
             comp=torch.zeros(T, dtype=torch.float32, device=dev),
             oinv=torch.zeros(T, 6, dtype=torch.float32, device=dev))
    for t in range(T):
        if kind == 'walking':
            obs, r, d, _ = env.step(acts[t])
        elif kind == 'flying':
            obs, r, d, _ = env.step({'movement': mv[t], 'camera': cam[t], 'inventory': inv[t], 'placement': pl[t]})
        else:
            obs, r, d, _ = env.step((bt[t], cam[t]))
        H['pose'][t] = env._pose[:, 0]
        H['ist'][t] = env._istate[:, 0]
        H['reward'][t] = r[0]
        H['done'][t] = env._done[0]
        H['grid'][t] = env.grid[0].reshape(-1)
        H['ogrid'][t] = obs['grid'][0].reshape(-1)
        H['apos'][t] = obs['agentPos'][0]
        H['comp'][t] = obs['compass'][0, 0]
        H['oinv'][t] = obs['inventory'][0]
    H = {k: v.cpu().numpy() for k, v in H.items()}
    nm = tr['name']
    # integers first: they localise a divergence best
    gold_grid = np.stack([G.grid_at(tr, t) for t in range(T)])
    bad = np.nonzero((H['grid'] != gold_grid).any(axis=1))[0]
    assert len(bad) == 0, f'{nm}: grid differs first at step {bad[0]}'
    assert np.array_equal(H['ogrid'], gold_grid.astype(np.int32)), nm
    inv16 = H['ist'][:, 0:3].copy().view(np.int16).reshape(T, 6).astype(np.int32)
    assert np.array_equal(inv16, tr['inventory']), nm
    assert np.array_equal(H['ist'][:, 3] & 0xFF, tr['active']), nm
    assert np.array_equal((H['ist'][:, 3] >> 8) & 0xFF, tr['tis']), nm
    assert np.array_equal(H['ist'][:, 4], tr['step_no']), nm
    assert np.array_equal(H['ist'][:, 6], tr['max_int']), nm
    assert np.array_equal(H['done'].astype(bool), tr['done']), nm
    assert np.array_equal(H['reward'], tr['reward']), nm
    gold_pose = np.concatenate([tr['pos'], tr['rot'], tr['dy'][:, None]], axis=1)
    if exact_pose:
        bad = np.nonzero((H['pose'] != gold_pose).any(axis=1))[0]
        assert len(bad) == 0, f'{nm}: pose differs first at step {bad[0]}: {H["pose"][bad[0]]} vs {gold_pose[bad[0]]}'
        assert np.array_equal(H['apos'], tr['agentPos']), nm
        assert np.array_equal(H['comp'], tr['compass'][:, 0]), nm
    else:
        assert np.abs(H['pose'] - gold_pose).max() <= POSE_TOL, nm
        assert np.abs(H['apos'] - tr['agentPos']).max() <= 1e-4, nm
    assert np.array_equal(H['oinv'], tr['obs_inventory']), nm
    return H


@pytest.mark.parametrize('tr', G.walking_traces(), ids=lambda t: t['name'])
def test_walking_golden(tr):
    _replay_gpu(tr, 'walking', exact_pose=tr['name'] != 'odd_rotation')


@pytest.mark.parametrize('tr', G.walking_traces()[-6:], ids=lambda t: t['name'])
def test_walking_golden_reward_every_step(tr):
    """Evaluating maximal_intersection on every step must not change any output."""
    _replay_gpu(tr, 'walking', reward_mode='every_step')


@pytest.mark.parametrize('tr', G.flying_traces(), ids=lambda t: t['name'])
def test_flying_golden(tr):
    _replay_gpu(tr, 'flying', exact_pose=False)


@pytest.mark.parametrize('tr', G.dict_traces(), ids=lambda t: t['name'])
def test_walking_dict_golden(tr):
    _replay_gpu(tr, 'dict', exact_pose=False)


@pytest.mark.parametrize('kind', ['walking', 'flying', 'dict'])
def test_golden_traces_overlapped_step(kind):
    """The reference-recorded traces of every action space through the overlapped step (k_widen next to the physics
    launch, then k_commit), forced for these single-env worlds, with every-step reward."""
    traces = {'walking': G.walking_traces()[-8:], 'flying': G.flying_traces(), 'dict': G.dict_traces()}[kind]
    for tr in traces:
        _replay_gpu(tr, kind, reward_mode='every_step', exact_pose=(kind == 'walking' and tr['name'] != 'odd_rotation'),
                    bits=128)


# ------------------------------------------------------------------------- reward kernel

def test_reward_goals():
    import gridworld_b200 as gw
    z = G.load('goals.npz')
    T = z['targets']
    g1 = np.stack([np.rot90(t, k=-1, axes=(1, 2)) for t in T]).copy()
    assert np.array_equal(gw.max_intersection(g1, T).cpu().numpy(), z['ka1'])
    assert np.array_equal(gw.max_intersection(T, T).cpu().numpy(), z['target_size'])
    for j in range(z['grids'].shape[1]):
        mi, arg = gw.max_intersection(z['grids'][:, j], T, return_argmax=True)
        assert np.array_equal(mi.cpu().numpy(), z['max_int'][:, j])
        assert np.array_equal(arg.cpu().numpy(), z['argmax'][:, j])


def test_reward_pairs_golden():
    import gridworld_b200 as gw
    z = G.load('reward_pairs.npz')
    for inv in (True, False):
        for hf in (True, False):
            sel = np.nonzero((z['invariant'] == inv) & (z['has_full'] == hf))[0]
            if len(sel) == 0:
                continue
            mi, arg = gw.max_intersection(z['grids'][sel], z['targets'][sel], z['fulls'][sel] if hf else None,
                                          invariant=inv, return_argmax=True)
            assert np.array_equal(mi.cpu().numpy(), z['max_int'][sel]), (inv, hf)
            assert np.array_equal(arg.cpu().numpy(), z['argmax'][sel]), (inv, hf)


def _synth(rng, n, rho, support, classes=6):
    u = rng.random_sample((n, 9, 11, 11))
    t = np.where(u < rho, rng.randint(1, classes + 1, (n, 9, 11, 11)), 0).astype(np.int8)
    if support == 'compact':
        m = np.zeros((9, 11, 11), bool)
        m[:5, 3:8, 3:8] = True
        t = np.where(m[None], t, 0).astype(np.int8)
    return t


@pytest.mark.parametrize('rho', [0.05, 0.2, 0.5])
@pytest.mark.parametrize('support', ['full', 'compact'])
def test_reward_vs_oracle_random(rho, support):
    """config-4 style sweep at a size the C oracle finishes in seconds."""
    import gridworld_b200 as gw
    rng = np.random.RandomState(int(rho * 100) + (7 if support == 'full' else 13))
    n = 3000
    T = _synth(rng, n, rho, support)
    Gd = _synth(rng, n, rho, support)
    # half of the grids: the target itself rotated by k and with 10% noise (high scores)
    for i in range(0, n, 2):
        g = np.rot90(T[i], k=-rng.randint(4), axes=(1, 2)).copy()
        noise = rng.random_sample(g.shape) < 0.1
        Gd[i] = np.where(noise, Gd[i], g)
    want, warg = O.max_intersection(Gd, T, threads=8, return_argmax=True)
    got, garg = gw.max_intersection(Gd, T, return_argmax=True)
    assert np.array_equal(got.cpu().numpy(), want)
    assert np.array_equal(garg.cpu().numpy(), warg)
    # difference-grid classes (-6..6) as produced by env.py:230/290
    S = _synth(rng, n, rho / 2, support)
    Td = (T.astype(np.int16) - S).astype(np.int8)
    Gdd = (Gd.astype(np.int16) - S).astype(np.int8)
    want = O.max_intersection(Gdd, Td, threads=8)
    got = gw.max_intersection(Gdd, Td)
    assert np.array_equal(got.cpu().numpy(), want)
    # full_grid rectangles (clipping path) and invariant=False
    F = np.where(T != 0, T, _synth(rng, n, 0.02, 'full')).astype(np.int8)
    want = O.max_intersection(Gd[:500], T[:500], F[:500], threads=8)
    got = gw.max_intersection(Gd[:500], T[:500], F[:500])
    assert np.array_equal(got.cpu().numpy(), want)
    Z = np.zeros_like(T[:200])
    want = O.max_intersection(Gd[:200], T[:200], Z, threads=8)
    got = gw.max_intersection(Gd[:200], T[:200], Z)
    assert np.array_equal(got.cpu().numpy(), want)
    want = O.max_intersection(Gd[:500], T[:500], invariant=False, threads=8)
    got = gw.max_intersection(Gd[:500], T[:500], invariant=False)
    assert np.array_equal(got.cpu().numpy(), want)


def test_reward_zone_spanning_rows_with_values_outside_colour_range():
    """Zone-spanning crowded targets take the packed-column comparison only when both rows hold colour differences
    (-6..6) and the byte-wise one otherwise. Values outside that range match nothing here; with disjoint
    stray values in grid and target the reference's raw-integer comparison (the oracle) gives the same answer."""
    import gridworld_b200 as gw
    rng = np.random.RandomState(77)
    n = 1200
    T = _synth(rng, n, 0.3, 'full')
    G = _synth(rng, n, 0.3, 'full')
    for i in range(0, n, 2):
        G[i] = np.rot90(T[i], k=-rng.randint(4), axes=(1, 2))
    S = _synth(rng, n, 0.1, 'full')
    T = (T.astype(np.int16) - S).astype(np.int8)       # difference grids: negative classes too
    G = (G.astype(np.int16) - S).astype(np.int8)
    want0, warg0 = O.max_intersection(G, T, threads=8, return_argmax=True)
    got0, garg0 = gw.max_intersection(G, T, return_argmax=True)   # all in range: packed columns
    assert np.array_equal(got0.cpu().numpy(), want0) and np.array_equal(garg0.cpu().numpy(), warg0)
    Gs, Ts = G.copy(), T.copy()
    for i in range(n):                                   # every third pair stays in range
        if i % 3 == 0:
            Gs[i].reshape(-1)[rng.randint(1089, size=3)] = rng.choice([100, 17, 7, -7, -128])
        if i % 3 == 1:
            Ts[i].reshape(-1)[rng.randint(1089, size=3)] = rng.choice([-100, -17, 8, -8, 127])
    want, warg = O.max_intersection(Gs, Ts, threads=8, return_argmax=True)
    got, garg = gw.max_intersection(Gs, Ts, return_argmax=True)
    assert np.array_equal(got.cpu().numpy(), want)
    assert np.array_equal(garg.cpu().numpy(), warg)
    got = gw.max_intersection(Gs, Ts, invariant=False)
    assert np.array_equal(got.cpu().numpy(), O.max_intersection(Gs, Ts, invariant=False, threads=8))


def test_reward_properties_full_size():
    """1M pairs (BASELINE config 4 size): size-independent properties of maximal_intersection."""
    import gridworld_b200 as gw
    n = 1_000_000
    gen = torch.Generator(device='cuda').manual_seed(1234)
    u = torch.rand(n, 9, 11, 11, device='cuda', generator=gen)
    col = torch.randint(1, 7, (n, 9, 11, 11), device='cuda', generator=gen, dtype=torch.int8)
    T = torch.where(u < 0.2, col, torch.zeros_like(col))
    size = (T != 0).flatten(1).sum(1).to(torch.int32)
    # (1) a grid equal to the target scores target_size; (2) so does any y-rotation of it
    assert torch.equal(gw.max_intersection(T, T), size)
    R = torch.rot90(T, k=1, dims=(2, 3)).contiguous()
    assert torch.equal(gw.max_intersection(R, T), size)
    # (3) bounded by target_size and by the grid's block count; (4) empty grid scores 0
    Gd = torch.where(torch.rand(n, 9, 11, 11, device='cuda', generator=gen) < 0.2, col.flip(0), torch.zeros_like(col))
    mi = gw.max_intersection(Gd, T)
    assert bool((mi <= size).all()) and bool((mi >= 0).all())
    assert bool((mi <= (Gd != 0).flatten(1).sum(1)).all())
    assert int(gw.max_intersection(torch.zeros_like(T[:1000]), T[:1000]).abs().sum()) == 0
    # (7) a target that spans the zone in x and z admits the shift (0, 0) only (task.py:67-72): the score is the best of
    #     the four rotations' cell-wise agreement, which torch can count directly -- an exact check of all 1 M pairs
    occ = (T != 0).any(dim=1)                                   # [n, x, z]
    spans = occ.any(dim=2)[:, 0] & occ.any(dim=2)[:, 10] & occ.any(dim=1)[:, 0] & occ.any(dim=1)[:, 10]
    assert int(spans.sum()) > n - 10
    for G in (Gd, R):
        want = torch.zeros(n, dtype=torch.int32, device='cuda')
        for k in range(4):
            Tk = torch.rot90(T, k=k, dims=(2, 3))
            want = torch.maximum(want, ((Tk != 0) & (Tk == G)).flatten(1).sum(1).to(torch.int32))
        got = gw.max_intersection(G, T)
        assert torch.equal(got[spans], want[spans])
    del occ, want, got
    # (5) the score only depends on the target up to an in-zone XZ translation (the admissible shifts
    #     enumerate every in-zone translate of every rotation): move a compact target, same result
    m = torch.zeros(9, 11, 11, dtype=torch.bool, device='cuda')
    m[:5, 3:8, 3:8] = True
    Tc = torch.where(m, T, torch.zeros_like(T))[:200_000]
    Gc = Gd[:200_000]
    base = gw.max_intersection(Gc, Tc)
    moved = torch.roll(Tc, shifts=(2, -3), dims=(2, 3))  # support stays inside the zone: no wrap
    assert torch.equal(gw.max_intersection(Gc, moved), base)
    # (6) ... and up to a y-rotation of the target
    assert torch.equal(gw.max_intersection(Gc, torch.rot90(Tc, k=1, dims=(2, 3)).contiguous()), base)


# ------------------------------------------------------------------------- batched env vs the C oracle

def _tasks(rng, n):
    T = _synth(rng, n, 0.12, 'compact')
    S = _synth(rng, n, 0.04, 'compact')
    S[:, 2:] = 0
    S[:, :, 5, 5] = 0
    S[::3] = 0  # a third of the envs start empty
    poses = np.zeros((n, 5))
    poses[::5, 0] = rng.randint(-8, 9, len(poses[::5])) * 0.25
    poses[::5, 2] = rng.randint(-8, 9, len(poses[::5])) * 0.25
    poses[::5, 3] = rng.randint(0, 72, len(poses[::5])) * 5.0
    poses[::7, 4] = rng.randint(-18, 19, len(poses[::7])) * 5.0
    return T, S, poses


# scheduling switches of iglu_config.reserved (results identical): 128 = the overlapped step (k_widen next to the physics
# launch, then k_commit) also for small batches, 64 = never; 32 = group-wise instead of tiled observe launch
OVERLAP, NO_OVERLAP, GROUPWISE = 128, 64, 32
SMALL_COMMIT = 32768   # 3-CTA commit launch: every warp works through a dozen records (its two-slot ring)


@pytest.mark.parametrize('reward_mode,alternate,grid_obs,bits', [('on_change', True, True, 0), ('every_step', True, True, 0),
                                                                 ('every_step', False, True, 0),
                                                                 ('on_change', False, 'incremental', 0),
                                                                 ('every_step', True, 'incremental', 0),
                                                                 ('every_step', True, 'incremental', GROUPWISE),
                                                                 ('every_step', True, False, 0),
                                                                 ('on_change', True, False, 0),
                                                                 ('on_change', True, False, NO_OVERLAP),
                                                                 ('on_change', True, 'incremental', NO_OVERLAP),
                                                                 ('on_change', True, True, OVERLAP),
                                                                 ('every_step', True, True, OVERLAP),
                                                                 ('every_step', True, True, OVERLAP | SMALL_COMMIT)])
def test_walking_batch_vs_oracle(reward_mode, alternate, grid_obs, bits):
    import gridworld_b200 as gw
    n, steps = 2048, 400
    rng = np.random.RandomState(3)
    T, S, poses = _tasks(rng, n)
    env = gw.make_vec(n, action_space='walking', max_steps=100, auto_reset=True, reward_mode=reward_mode,
                      alternate_phases=alternate, grid_obs=grid_obs, _sched_bits=bits)
    env.set_tasks(T, S, poses)
    orc = O.OracleBatch(n, 'walking', max_steps=100, threads=8)
    orc.set_tasks(T, S, poses)
    p = np.array([1, 5, 2, 2, 2, 3, 2, 2, 2, 2, 2, 2, 2, 2, 5, 2, 4, 6], float)
    for t in range(steps):
        a = rng.choice(18, n, p=p / p.sum()).astype(np.int32)
        _, r, d, _ = env.step(a)
        ro, do = orc.step_walking(a, auto_reset=True)
        assert np.array_equal(r.cpu().numpy(), ro), t
        assert np.array_equal(d.cpu().numpy(), do), t
        if t % 25 == 24 or t == steps - 1:
            s = orc.state()
            assert np.array_equal(env.grid.cpu().numpy(), s['grid']), t
            assert np.array_equal(env.position.cpu().numpy(), s['pos']), t
            assert np.array_equal(env.rotation.cpu().numpy(), s['rot']), t
            assert np.array_equal(env.dy.cpu().numpy(), s['dy']), t
            assert np.array_equal(env.inventory.cpu().numpy(), s['inventory']), t
            assert np.array_equal(env.max_int.cpu().numpy(), s['max_int']), t
            assert np.array_equal(env.step_no.cpu().numpy(), s['step_no']), t
            assert np.array_equal(env.time_int_steps.cpu().numpy(), s['tis']), t
            o = orc.obs()
            ob = env.observation()
            if grid_obs:
                assert np.array_equal(ob['grid'].cpu().numpy(), o['grid']), t
            assert np.array_equal(ob['agentPos'].cpu().numpy(), o['agentPos']), t
            assert np.array_equal(ob['compass'].cpu().numpy(), o['compass']), t
            assert np.array_equal(ob['inventory'].cpu().numpy(), o['inventory']), t


@pytest.mark.parametrize('reward_mode,bits', [('on_change', 0), ('every_step', 0), ('every_step', OVERLAP)])
def test_walking_zone_spanning_targets_vs_oracle(reward_mode, bits):
    """Targets that span the zone in x and z admit one shift per rotation (task.py:57-72): the kernels then
    compare grid bytes directly instead of building occupancy rows. With and without starting grids."""
    import gridworld_b200 as gw
    n, steps = 1024, 300
    rng = np.random.RandomState(17)
    T = _synth(rng, n, 0.15, 'full')
    T[:, :, 0, :] = 1  # make sure x and z extents are 0..10 for every env
    T[:, :, 10, :] = 2
    T[:, :, :, 0] = 3
    T[:, :, :, 10] = 4
    T[:, 2:] = 0
    T[::4] = _synth(rng, len(T[::4]), 0.1, 'compact')  # a quarter keeps many shifts (row path in the same launch)
    _, S, poses = _tasks(rng, n)
    env = gw.make_vec(n, max_steps=90, auto_reset=True, reward_mode=reward_mode, _sched_bits=bits)
    env.set_tasks(T, S, poses)
    orc = O.OracleBatch(n, 'walking', max_steps=90, threads=8)
    orc.set_tasks(T, S, poses)
    p = np.array([1, 5, 2, 2, 2, 3, 2, 2, 2, 2, 2, 2, 2, 2, 5, 2, 4, 6], float)
    for t in range(steps):
        a = rng.choice(18, n, p=p / p.sum()).astype(np.int32)
        _, r, d, _ = env.step(a)
        ro, do = orc.step_walking(a, auto_reset=True)
        assert np.array_equal(r.cpu().numpy(), ro), t
        assert np.array_equal(d.cpu().numpy(), do), t
    s = orc.state()
    assert np.array_equal(env.grid.cpu().numpy(), s['grid'])
    assert np.array_equal(env.max_int.cpu().numpy(), s['max_int'])
    assert (s['max_int'] > 0).mean() > 0.1  # the scenario really scores


@pytest.mark.parametrize('bits', [0, OVERLAP])
def test_flying_batch_vs_oracle(bits):
    import gridworld_b200 as gw
    n, steps = 1024, 300
    rng = np.random.RandomState(5)
    T, S, poses = _tasks(rng, n)
    env = gw.make_vec(n, action_space='flying', max_steps=120, auto_reset=True, target_in_obs=True, _sched_bits=bits)
    env.set_tasks(T, S, poses)
    orc = O.OracleBatch(n, 'flying', max_steps=120, threads=8)
    orc.set_tasks(T, S, poses)
    bad_envs = np.zeros(n, bool)  # envs where a measure-zero rounding flip made the traces diverge
    for t in range(steps):
        mv = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
        mv[rng.random_sample((n, 3)) < 0.3] = 0
        cam = rng.uniform(-5, 5, (n, 2)).astype(np.float32)
        inv = rng.randint(0, 7, n).astype(np.int32)
        pl = rng.randint(0, 3, n).astype(np.int32)
        obs, r, d, _ = env.step({'movement': mv, 'camera': cam, 'inventory': inv, 'placement': pl})
        ro, do = orc.step_flying(mv, cam, inv, pl, auto_reset=True)
        s = orc.state()
        pos_err = np.abs(env.position.cpu().numpy() - s['pos']).max(axis=1)
        grid_bad = (env.grid.cpu().numpy() != s['grid']).reshape(n, -1).any(axis=1)
        bad_envs |= (pos_err > POSE_TOL) | grid_bad
        ok = ~bad_envs
        assert np.array_equal(r.cpu().numpy()[ok], ro[ok]), t
        assert np.array_equal(d.cpu().numpy()[ok], do[ok]), t
    # device sincos/atan2 differ from glibc by <= 2 ulp: trajectories agree to 1e-6 except where a
    # discontinuity (round-half, zone gate) flips; that must stay a vanishing fraction
    assert bad_envs.mean() <= 0.002, bad_envs.sum()
    assert np.array_equal(obs['target_grid'].cpu().numpy(), T.astype(np.int32))


def test_sharding_is_concatenation():
    """Multi-GPU layout = contiguous env blocks with no exchange: stepping two half-size envs must equal
    the full-size env (the indexing test of SURVEY 8(e)); uses one device."""
    import gridworld_b200 as gw
    n = 4096
    rng = np.random.RandomState(9)
    T, S, poses = _tasks(rng, n)
    full = gw.make_vec(n, max_steps=60, reward_mode='every_step')
    full.set_tasks(T, S, poses)
    h = n // 2
    parts = [gw.make_vec(h, max_steps=60, reward_mode='every_step') for _ in range(2)]
    parts[0].set_tasks(T[:h], S[:h], poses[:h])
    parts[1].set_tasks(T[h:], S[h:], poses[h:])
    for t in range(150):
        a = torch.as_tensor(rng.randint(0, 18, n).astype(np.int32), device='cuda')
        _, r, d, _ = full.step(a)
        r0 = parts[0].step(a[:h])[1]
        r1 = parts[1].step(a[h:])[1]
        assert torch.equal(r, torch.cat([r0, r1])), t
    assert torch.equal(full.grid, torch.cat([parts[0].grid, parts[1].grid]))
    assert torch.equal(full._pose, torch.cat([parts[0]._pose, parts[1]._pose], dim=1))


@pytest.mark.parametrize('tr', G.generator_traces(), ids=lambda t: t['name'])
def test_task_generator_golden(tr):
    """env.set_task_generator on the reference (every reset draws the next task, env.py:213-216) vs the
    device-resident pool: the indices the reference's generator drew are imposed through next_task, everything
    else -- Task.__init__ products, starting grid, inventory, target_grid obs -- comes from the pool on device."""
    import gridworld_b200 as gw
    flying = tr['space'] == 'flying'
    env = gw.make_vec(1, action_space=tr['space'], max_steps=tr['max_steps'], auto_reset=True, target_in_obs=True)
    dev = env.device
    ch = [int(c) for c in tr['choices']]
    nt = torch.full((1,), ch.pop(0), dtype=torch.int32, device=dev)
    env.set_task_pool(tr['pool_targets'], tr['pool_starts'], seed=11, next_task=nt)
    a = tr['actions']
    done = False
    for t in range(tr['n']):
        if done:
            nt[0] = ch.pop(0)
        if flying:
            obs, r, d, _ = env.step({'movement': a['movement'][t][None], 'camera': a['camera'][t][None],
                                     'inventory': a['inv'][t:t + 1], 'placement': a['placement'][t:t + 1]})
        else:
            obs, r, d, _ = env.step(a['actions'][t:t + 1])
        done = bool(d[0])
        ctx = f"{tr['name']} step {t}"
        j = int(tr['task_at_step'][t])
        assert int(env.task_ids[0]) == j and int(nt[0]) == -1, ctx
        assert float(r[0]) == tr['reward'][t], ctx
        assert done == bool(tr['done'][t]), ctx
        assert np.array_equal(env.grid[0].cpu().numpy().reshape(-1), G.grid_at(tr, t)), ctx
        assert np.array_equal(obs['grid'][0].cpu().numpy().reshape(-1), G.grid_at(tr, t).astype(np.int32)), ctx
        assert np.array_equal(obs['target_grid'][0].cpu().numpy(), tr['pool_targets'][j].astype(np.int32)), ctx
        assert env.inventory[0].cpu().numpy().tolist() == tr['inventory'][t].tolist(), ctx
        assert int(env.max_int[0]) == tr['max_int'][t] and int(env.step_no[0]) == tr['step_no'][t], ctx
        pos = env.position[0].cpu().numpy()
        if flying:
            assert np.abs(pos - tr['pos'][t]).max() <= POSE_TOL, ctx
        else:
            assert pos.tolist() == tr['pos'][t].tolist() and float(env.dy[0]) == tr['dy'][t], ctx
            assert np.array_equal(obs['agentPos'][0].cpu().numpy(), tr['agentPos'][t]), ctx
    assert not ch


def _pool(rng, m):
    T = _synth(rng, m, 0.1, 'compact')
    S = _synth(rng, m, 0.04, 'compact')
    S[:, 2:] = 0
    S[:, :, 5, 5] = 0
    S[::2] = 0
    T[0] = 0
    T[0, 0, 5, 4] = 1  # one-block task (finishes early when the block lands there)
    return T, S


@pytest.mark.parametrize('reward_mode,grid_obs,bits', [('on_change', True, 0), ('every_step', True, 0),
                                                       ('on_change', 'incremental', 0), ('every_step', True, OVERLAP),
                                                       ('on_change', True, OVERLAP)])
def test_task_generator_batch_vs_oracle(reward_mode, grid_obs, bits):
    """Counter-based draws on device (fused and two-launch step forms) vs the oracle's restatement."""
    import gridworld_b200 as gw
    n, m, steps = 1536, 37, 260
    rng = np.random.RandomState(21)
    T, S = _pool(rng, m)
    env = gw.make_vec(n, max_steps=24, auto_reset=True, reward_mode=reward_mode, grid_obs=grid_obs, target_in_obs=True,
                      _sched_bits=bits)
    env.set_task_pool(T, S, seed=0x1234567890ABCDEF, env_offset=5000)
    orc = O.OracleBatch(n, 'walking', max_steps=24, threads=8)
    orc.set_pool(T, S, seed=0x1234567890ABCDEF, env_offset=5000)
    orc.reset()
    assert np.array_equal(env.task_ids.cpu().numpy(), orc.task_ids)
    p = np.array([1, 5, 2, 2, 2, 3, 2, 2, 2, 2, 2, 2, 2, 2, 5, 2, 4, 6], float)
    for t in range(steps):
        a = rng.choice(18, n, p=p / p.sum()).astype(np.int32)
        if t == 100:  # explicit masked reset in the middle: draws for the masked envs only
            mask = (np.arange(n) % 3 == 0).astype(np.uint8)
            env.reset(mask)
            orc.reset(mask)
        obs, r, d, _ = env.step(a)
        ro, do = orc.step_walking(a, auto_reset=True)
        assert np.array_equal(r.cpu().numpy(), ro), t
        assert np.array_equal(d.cpu().numpy(), do), t
        assert np.array_equal(env.task_ids.cpu().numpy(), orc.task_ids), t
        if t % 20 == 19 or t == steps - 1:
            s = orc.state()
            assert np.array_equal(env.grid.cpu().numpy(), s['grid']), t
            assert np.array_equal(env.position.cpu().numpy(), s['pos']), t
            assert np.array_equal(env.inventory.cpu().numpy(), s['inventory']), t
            assert np.array_equal(env.max_int.cpu().numpy(), s['max_int']), t
            assert np.array_equal(obs['grid'].cpu().numpy(), orc.obs()['grid']), t
            assert np.array_equal(obs['target_grid'].cpu().numpy(), orc.target_grids()), t
            assert np.array_equal(obs['target_grid'].cpu().numpy(), T[orc.task_ids].astype(np.int32)), t
    assert len(np.unique(orc.task_ids)) > m // 2


def test_set_task_generator_from_task_collection():
    """set_task_generator(collection): the collection is enumerated into the device pool; targets / starting grids /
    dialogs of the drawn tasks are what the host objects say."""
    import gridworld_b200 as gw
    z = G.load('goals.npz')
    names = [str(nm) for nm in z['names'][:20]]
    start = [(0, -1, 1, 3), (1, -1, 1, 4)]
    tasks = [gw.Task(nm, z['targets'][i].astype(np.int32), starting_grid=start if i % 3 == 0 else [])
             for i, nm in enumerate(names)]
    n = 256
    env = gw.make_vec(n, max_steps=5, target_in_obs=True)
    obs = env.set_task_generator(tasks, seed=3)
    T, S, chats = gw.pool_arrays(tasks)
    for t in range(12):  # two auto-resets per env
        ids = env.task_ids.cpu().numpy()
        assert np.array_equal(obs['target_grid'].cpu().numpy(), T[ids].astype(np.int32))
        assert np.array_equal(env.starting_grid.cpu().numpy(), S[ids])
        assert [env.dialogs[j] for j in ids[:5]] == [names[j] for j in ids[:5]]
        obs, _, _, _ = env.step(np.zeros(n, np.int32))
    ids = env.task_ids.cpu().numpy()
    assert len(np.unique(ids)) > 10
    inv = env.inventory.cpu().numpy()  # envs whose task has the 2-block starting grid hold 19 of colours 3 and 4
    has = (ids % 3 == 0)
    assert (inv[has][:, 2] == 19).all() and (inv[has][:, 3] == 19).all() and (inv[~has] == 20).all()


def test_task_generator_sharding_and_checkpoint():
    """env_offset makes shards draw what the single-device run draws; state_dict carries the cursor."""
    import gridworld_b200 as gw
    n, m, h = 1024, 19, 512
    rng = np.random.RandomState(23)
    T, S = _pool(rng, m)
    full = gw.make_vec(n, max_steps=15)
    full.set_task_pool(T, S, seed=77)
    parts = [gw.make_vec(h, max_steps=15) for _ in range(2)]
    parts[0].set_task_pool(T, S, seed=77, env_offset=0)
    parts[1].set_task_pool(T, S, seed=77, env_offset=h)
    acts = [torch.as_tensor(rng.randint(0, 18, n).astype(np.int32), device='cuda') for _ in range(70)]
    for t, a in enumerate(acts):
        if t == 40:
            sd = full.state_dict()
        _, r, d, _ = full.step(a)
        r0 = parts[0].step(a[:h])[1]
        r1 = parts[1].step(a[h:])[1]
        assert torch.equal(r, torch.cat([r0, r1])), t
        assert torch.equal(full.task_ids, torch.cat([parts[0].task_ids, parts[1].task_ids])), t
    assert torch.equal(full.grid, torch.cat([parts[0].grid, parts[1].grid]))
    # resume from the step-40 checkpoint in a fresh env: same continuation
    again = gw.make_vec(n, max_steps=15)
    again.set_task_pool(T, S, seed=77, reset=False)
    again.load_state_dict(sd)
    for a in acts[40:]:
        again.step(a)
    assert torch.equal(again.grid, full.grid) and torch.equal(again.task_ids, full.task_ids)
    assert torch.equal(again._istate, full._istate)


def test_obs_grid_negative_cells_and_all_alignments():
    """obs['grid'] is the sign-extended int8 state for every value and every 16-byte phase of the row
    (WHITE = -1 is a legal value of the observation space, env.py:88)."""
    import gridworld_b200 as gw
    n = 67
    rng = np.random.RandomState(4)
    for kw in (dict(reward_mode='every_step'), dict(reward_mode='on_change')):
        env = gw.make_vec(n, max_steps=1000, **kw)
        env.set_tasks(_synth(rng, n, 0.1, 'compact'))
        g = rng.randint(-6, 7, (n, 1089)).astype(np.int8)
        g[rng.random_sample((n, 1089)) < 0.5] = 0
        g.reshape(n, 9, 11, 11)[:, 0:4, 4:7, 4:7] = 0  # free space around the agent
        env._grid[:, :1089] = torch.as_tensor(g, device=env.device)
        obs, _, _, _ = env.step(np.zeros(n, np.int32))
        assert np.array_equal(obs['grid'].cpu().numpy().reshape(n, 1089), g.astype(np.int32))


def test_step_host_routes():
    """iglu_step_walking_host: pinned buffers (kernels read / write host memory over PCIe), pageable buffers and
    the forced staged-copy route all equal the device-tensor step."""
    import gridworld_b200 as gw
    n, steps = 3000, 120
    rng = np.random.RandomState(31)
    T, S, poses = _tasks(rng, n)
    ref = gw.make_vec(n, max_steps=50, reward_mode='every_step')
    ref.set_tasks(T, S, poses)
    variants = []
    for pinned, bits in ((True, 0), (False, 0), (True, 16), (True, OVERLAP), (False, OVERLAP)):   # + the overlapped step
        env = gw.make_vec(n, max_steps=50, reward_mode='every_step', _sched_bits=bits)
        env.set_tasks(T, S, poses)
        mk = (lambda t: t.pin_memory()) if pinned else (lambda t: t)
        variants.append((env, mk(torch.zeros(n, dtype=torch.int32)), mk(torch.zeros(n, dtype=torch.float64)),
                         mk(torch.zeros(n, dtype=torch.uint8))))
    for t in range(steps):
        a = torch.as_tensor(rng.randint(0, 18, n).astype(np.int32))
        _, r, d, _ = ref.step(a.cuda())
        r, d = r.cpu(), d.cpu().to(torch.uint8)
        for env, ah, rh, dh in variants:
            ah.copy_(a)
            env.step_host(ah, rh, dh)
            assert torch.equal(rh, r) and torch.equal(dh, d), t
            assert torch.equal(env._reward.cpu(), r), t  # the device-side result buffers are written as well
    for env, *_ in variants:
        assert torch.equal(env.grid, ref.grid) and torch.equal(env._pose, ref._pose)


@pytest.mark.parametrize('reward_mode,bits', [('on_change', 0), ('every_step', 0), ('every_step', OVERLAP)])
def test_step_is_cuda_graph_capturable(reward_mode, bits):
    """The step entry points only enqueue work on the caller's stream (no allocation, no sync), so a sequence of
    steps can be captured once and replayed; replays equal eager stepping."""
    import gridworld_b200 as gw
    n, per = 700, 6
    rng = np.random.RandomState(41)
    T, S, poses = _tasks(rng, n)
    eager = gw.make_vec(n, max_steps=20, reward_mode=reward_mode, _sched_bits=bits)
    eager.set_tasks(T, S, poses)
    graphed = gw.make_vec(n, max_steps=20, reward_mode=reward_mode, _sched_bits=bits)
    graphed.set_tasks(T, S, poses)
    acts = torch.zeros(per, n, dtype=torch.int32, device='cuda')  # static input buffer of the graph
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):  # warm-up on a side stream (lazy module loading must not happen during capture)
        warm = gw.make_vec(n, max_steps=20, reward_mode=reward_mode, _sched_bits=bits)
        warm.set_tasks(T, S, poses)
        warm.step(acts[0])
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    rsum = torch.zeros(n, dtype=torch.float64, device='cuda')
    with torch.cuda.graph(g):
        for k in range(per):
            _, r, _, _ = graphed.step(acts[k])
            rsum += r
    # the capture itself does not execute anything: state is still the reset state
    want = torch.zeros(n, dtype=torch.float64, device='cuda')
    for rep in range(7):
        a = torch.as_tensor(rng.randint(0, 18, (per, n)).astype(np.int32), device='cuda')
        acts.copy_(a)
        g.replay()
        for k in range(per):
            _, r, _, _ = eager.step(a[k])
            want += r
    torch.cuda.synchronize()
    assert torch.equal(rsum, want)
    assert torch.equal(graphed.grid, eager.grid) and torch.equal(graphed._pose, eager._pose)
    assert torch.equal(graphed._istate, eager._istate)


def test_block_list_capacity_is_enforced():
    import gridworld_b200 as gw
    rng = np.random.RandomState(2)
    env = gw.make_vec(8, max_target_blocks=64)
    env.set_tasks(_synth(rng, 8, 0.03, 'full'))          # ~33 blocks: fits
    with pytest.raises(ValueError, match='max_target_blocks'):
        env.set_tasks(_synth(rng, 8, 0.3, 'full'))       # ~330 blocks: refused, not truncated
    with pytest.raises(ValueError, match='max_target_blocks'):
        env.set_task_pool(_synth(rng, 5, 0.3, 'full'))


def test_full_size_invariants():
    """BASELINE config 2 size (65,536 envs): properties that need no oracle."""
    import gridworld_b200 as gw
    n = 65536
    rng = np.random.RandomState(1)
    T = _synth(rng, n, 0.2, 'compact')
    env = gw.make_vec(n, max_steps=250, reward_mode='every_step')
    env.set_tasks(T)
    dev = env.device
    total_r = torch.zeros(n, dtype=torch.float64, device=dev)
    for t in range(300):
        a = torch.randint(0, 18, (n,), device=dev, dtype=torch.int32)
        obs, r, d, _ = env.step(a)
        total_r += r
        if t % 50 == 49:
            g = env.grid
            nblk = (g != 0).flatten(1).sum(1)
            inv = env.inventory
            # blocks in the world + blocks in the inventory is conserved (start empty: 120)
            assert bool(((inv.sum(1) + nblk) == 120).all())
            # per colour too
            for c in range(1, 7):
                assert bool((((g == c).flatten(1).sum(1)) + inv[:, c - 1] == 20).all())
            assert torch.equal(obs['grid'], g.to(torch.int32))
            assert bool((env.max_int <= env.target_size).all())
            # cached max_int equals a fresh stateless evaluation (empty start: never stale)
            assert torch.equal(gw.max_intersection(g, T).to(torch.int32), env.max_int)
            p = env.position
            assert bool((p[:, 0].abs() <= 7.25).all() and (p[:, 2].abs() <= 7.25).all())
            assert bool((env.step_no >= 1).all() and (env.step_no <= 250).all())
    assert int(env.error_flags.sum()) == 0


def test_dropin_gym_surface():
    """gym.make('IGLUGridworld-v0', ...) surface at N = 1 against the golden known-answer trace."""
    import gridworld_b200 as gw
    tr = [t for t in G.walking_traces() if t['name'] == 'ka_physics'][0]
    env = gw.make('IGLUGridworld-v0', render=False, vector_state=True, size_reward=False,
                  select_and_place=False, action_space='walking')
    with pytest.raises(ValueError):
        env.step(0)
    sparse = [(0, -1, -2, 1), (0, 0, -2, 1)]
    env.set_task(gw.Task('', gw.Tasks.to_dense(sparse), starting_grid=[]))
    obs = env.reset()
    assert obs['agentPos'].tolist() == [0, 0, 0, 0, 0] and obs['grid'].dtype == np.int32
    assert obs['grid'].shape == (9, 11, 11) and obs['inventory'].tolist() == [20.] * 6
    for t, a in enumerate(tr['actions']['actions']):
        obs, r, done, info = env.step(int(a))
        assert r == tr['reward'][t] and done == bool(tr['done'][t]) and info == {}
        assert list(env.agent.position) == tr['pos'][t].tolist()
        assert np.array_equal(obs['agentPos'], tr['agentPos'][t])
        assert np.array_equal(obs['grid'].reshape(-1), G.grid_at(tr, t).astype(np.int32))
    assert env.action_space.n == 18
    assert set(env.observation_space.spaces) == {'inventory', 'compass', 'dialog', 'agentPos', 'grid'}
    # flying + target_in_obs surface
    fenv = gw.make('IGLUGridworld-v0', render=False, vector_state=True, size_reward=True, action_space='flying',
                   target_in_obs=True)
    fenv.set_task(gw.Task('hello', gw.Tasks.to_dense(sparse), starting_grid=[(1, -1, 1, 3)]))
    obs = fenv.reset()
    assert obs['dialog'] == 'hello' and obs['inventory'].tolist() == [20, 20, 19, 20, 20, 20]
    assert obs['target_grid'][0, 5, 3] == 1 and obs['grid'][0, 6, 6] == 3
    obs, r, done, _ = fenv.step({'movement': np.array([1, 0, 0], np.float32), 'camera': np.zeros(2, np.float32),
                                 'inventory': 0, 'placement': 0})
    assert abs(fenv.agent.position[2] - 0.75) < 1e-12  # strafe[0] > 0 moves +z at 15 units/s


def test_get_intersection_and_subtasks_golden():
    """'next' rows: Task.get_intersection against the reference's outputs; the Actions wrapper."""
    import gridworld_b200 as gw
    z = G.load('misc.npz')
    task = gw.Task('', z['gi_target'].astype(np.int32))
    got = [task.get_intersection(z['gi_grid'], int(dx), int(dz), int(r)) for dx, dz, r in z['gi_query']]
    assert got == z['gi_value'].tolist()
    # Actions wrapper (wrappers.py:11-32): 17 actions, 'place' removed
    env = gw.Actions(gw.make('IGLUGridworld-v0', render=False, vector_state=True, size_reward=False))
    assert env.action_space.n == 17
    env.set_task(gw.Task('', gw.Tasks.to_dense([(0, -1, -2, 1)]), starting_grid=[]))
    obs, r, done, _ = env.step(16)
    assert r == 0 and not done


def test_new_host_surface_async_snapshot_set_grids_restore():
    """Round-2 additions of the boundary: iglu_step_walking_host_async / iglu_wait (two groups in flight give the
    same results as one synchronous env), iglu_fetch_env_host, IGLUVecEnv.set_grids, the device evaluation
    counter, load_state_dict rebuilding the incrementally maintained observation buffers, and the refusal to step
    an env that was never reset."""
    import gridworld_b200 as gw
    from bench import shard_state, synth_targets
    n, steps = 512, 40
    T = synth_targets(n, 0.2, 'compact', 5)
    rng = np.random.RandomState(1)
    acts = torch.as_tensor(rng.randint(0, 18, (steps, n)).astype(np.int32))
    ref = gw.make_vec(n, max_steps=30, reward_mode='every_step')
    with pytest.raises(ValueError):
        ref.step(acts[0].cuda())                       # no task yet
    ref.set_tasks(T, reset=False)
    with pytest.raises(ValueError, match='reset'):
        ref.step(acts[0].cuda())                       # task set but never reset (reference env.py:275)
    ref.reset()
    sd0 = ref.state_dict()
    R = np.zeros((steps, n))
    Dn = np.zeros((steps, n), bool)
    ref.reset_evaluation_counter()
    nonempty = 0
    for t in range(steps):
        _, r, d, _ = ref.step(acts[t].cuda())
        R[t], Dn[t] = r.cpu().numpy(), d.cpu().numpy()
        nonempty += int((ref._istate[5] != 0).sum())
    assert ref.evaluations == nonempty                  # every env with a non-empty difference grid was evaluated
    # two half-size envs stepped through the async host call on their own streams, results consumed after wait()
    halves = []
    for lo, hi in ((0, n // 2), (n // 2, n)):
        e = gw.make_vec(hi - lo, max_steps=30, reward_mode='every_step')
        e.load_state_dict(shard_state(sd0, lo, hi))
        halves.append((e, torch.cuda.Stream(), acts[:, lo:hi].contiguous().pin_memory(),
                       torch.zeros(hi - lo, dtype=torch.float64).pin_memory(),
                       torch.zeros(hi - lo, dtype=torch.uint8).pin_memory(), lo, hi))
    for t in range(steps):
        for e, s, a, r, d, lo, hi in halves:
            with torch.cuda.stream(s):
                e.step_host(a[t], r, d, wait=False)
        for e, s, a, r, d, lo, hi in halves:
            e.wait()
            assert np.array_equal(r.numpy(), R[t, lo:hi]) and np.array_equal(d.numpy().astype(bool), Dn[t, lo:hi]), t
    torch.cuda.synchronize()
    for e, s, a, r, d, lo, hi in halves:
        assert torch.equal(e.grid, ref.grid[lo:hi]) and torch.equal(e._pose, ref._pose[:, lo:hi])
    # snapshot of one env == the batched tensors
    sn = ref.fetch_env(7)
    assert list(sn.pose) == ref._pose[:, 7].tolist() and sn.reward == float(ref._reward[7])
    assert np.array_equal(np.ctypeslib.as_array(sn.grid), ref.grid[7].reshape(-1).cpu().numpy())
    assert np.array_equal(np.ctypeslib.as_array(sn.target_grid), ref.target_grid[7].reshape(-1).cpu().numpy())
    assert list(sn.inventory) == ref.inventory[7].tolist() and sn.step_no == int(ref.step_no[7])
    assert sn.max_int == int(ref.max_int[7]) and sn.target_size == int(ref.target_size[7])
    assert list(sn.agent_pos) == ref._obs_pos[7].tolist() and sn.done == int(ref._done[7])
    # restore into envs that keep obs['grid'] incrementally / expose target_grid: derived buffers are rebuilt
    inc = gw.make_vec(n, max_steps=30, reward_mode='every_step', grid_obs='incremental', target_in_obs=True)
    inc.load_state_dict(ref.state_dict())
    assert torch.equal(inc.observation()['grid'], ref.grid.to(torch.int32))
    assert torch.equal(inc.observation()['target_grid'], ref.target_grid.to(torch.int32))
    for t in range(5):
        a = acts[t].cuda()
        o1, r1, d1, _ = inc.step(a)
        o2, r2, d2, _ = ref.step(a)
        assert torch.equal(o1['grid'], o2['grid']) and torch.equal(r1, r2) and torch.equal(d1, d2)
    # set_grids: pre-built worlds keep every invariant of a world the agent built itself
    env = gw.make_vec(n, max_steps=10 ** 6, reward_mode='every_step')
    env.set_tasks(T)
    built = torch.where(torch.rand(T.shape) < 0.5, T, torch.zeros((), dtype=torch.int8))
    built[:, :3, 5, 5] = 0
    env.set_grids(built)
    assert torch.equal(env.max_int, gw.max_intersection(built, T).to(torch.int32))
    assert bool((env.max_int == (built != 0).flatten(1).sum(1).cuda()).all())  # a subset of the target matches itself
    env.reset_evaluation_counter()
    for t in range(20):
        obs, r, d, _ = env.step(acts[t].cuda())
    assert env.evaluations == 20 * n                     # dense worlds: every env evaluated on every step
    g = env.grid
    for c in range(1, 7):
        assert bool((((g == c).flatten(1).sum(1)) + env.inventory[:, c - 1] == 20).all())
    assert torch.equal(gw.max_intersection(g, T).to(torch.int32), env.max_int)
    assert torch.equal(obs['grid'], g.to(torch.int32))
    # the env's own rows (stride 1104, 16-byte aligned) evaluated where they lie == the dense repack
    v1, a1 = gw.max_intersection(env._grid, env._target, return_argmax=True)
    v2, a2 = gw.max_intersection(g, T, return_argmax=True)
    assert torch.equal(v1, v2) and torch.equal(a1, a2)


def _sparse(d):
    ys, xs, zs = np.nonzero(d)
    return [(int(x) - 5, int(y) - 1, int(z) - 5, int(d[y, x, z])) for y, x, z in zip(ys, xs, zs)]


def test_surface_golden_size_reward_initialize_world_host_generator():
    """SURVEY rows a20 / a23 at N = 1 against traces recorded from the executed reference (tests/golden/surface.npz):
    the SizeReward wrapper's (degenerate) reward, initialize_world / deinitialize_world, and set_task_generator with
    a host-side generator object."""
    import gridworld_b200 as gw
    z = G.load('surface.npz')

    def check(env, t, pre, r, done):
        u = env.unwrapped
        assert r == z[pre + 'reward'][t] and done == bool(z[pre + 'done'][t]), (pre, t, r, z[pre + 'reward'][t])
        assert list(u.agent.position) + list(u.agent.rotation) == z[pre + 'pose'][t].tolist(), (pre, t)
        assert u.agent.inventory == z[pre + 'inv'][t].tolist(), (pre, t)
        assert np.array_equal(u.grid.reshape(-1), z[pre + 'grid'][t].astype(np.int32)), (pre, t)

    # ---- SizeReward (env.py:316-331) ----
    env = gw.make('IGLUGridworld-v0', render=False, vector_state=True, size_reward=True, action_space='walking',
                  max_steps=40)
    env.set_task(gw.Task('size', z['size_target'].astype(np.int32), starting_grid=_sparse(z['size_start'])))
    done = False
    for t, a in enumerate(z['size_actions']):
        if done:
            env.reset()
        assert done == bool(z['size_resets'][t])
        obs, r, done, _ = env.step(int(a))
        check(env, t, 'size_', r, done)
    assert env.unwrapped.max_int == int(z['size_max_int_attr'])
    assert 1.0 in z['size_reward'] and set(z['size_reward'].tolist()) <= {0.0, 1.0}

    # ---- initialize_world / deinitialize_world (env.py:177-193) ----
    env = gw.make('IGLUGridworld-v0', render=False, vector_state=True, size_reward=False, action_space='walking',
                  max_steps=60)
    env.set_task(gw.Task('init', z['iw_target'].astype(np.int32), starting_grid=_sparse(z['iw_start'])))
    events = {int(t): int(k) for t, k in z['iw_events']}
    done = False
    for t, a in enumerate(z['iw_actions']):
        ev = events.get(t, 0)
        if ev == 1:
            with pytest.warns(UserWarning):
                env.initialize_world(_sparse(z['iw_world']), tuple(z['iw_pose0'].tolist()))
            done = False
        elif ev == 2:
            env.deinitialize_world()
            done = False
        elif done:
            env.reset()
        obs, r, done, _ = env.step(int(a))
        check(env, t, 'iw_', r, done)
        assert np.array_equal(obs['agentPos'], z['iw_agentPos'][t])

    # ---- set_task_generator with a host-side generator (env.py:168-175, 213-216) ----
    class Cycle:
        def __init__(self, tasks):
            self.tasks, self.i = tasks, 0

        def reset(self):
            self.i += 1
            return self.tasks[(self.i - 1) % len(self.tasks)]

    tasks = [gw.Task(f'task {j}', z['gen_targets'][j].astype(np.int32), starting_grid=_sparse(z['gen_starts'][j]))
             for j in range(len(z['gen_targets']))]
    env = gw.make('IGLUGridworld-v0', render=False, vector_state=True, size_reward=False, action_space='walking',
                  max_steps=25, target_in_obs=True)
    env.set_task_generator(Cycle(tasks))
    done = False
    for t, a in enumerate(z['gen_actions']):
        if done:
            env.reset()
        assert done == bool(z['gen_resets'][t])
        obs, r, done, _ = env.step(int(a))
        check(env, t, 'gen_', r, done)
        assert obs['dialog'] == str(z['gen_dialog'][t])
        assert np.array_equal(obs['target_grid'].reshape(-1), z['gen_target_obs'][t])
    with pytest.warns(UserWarning):
        env.set_task(tasks[0])   # env.py:162-164: no effect on the generator, but warns


@pytest.mark.parametrize('support', ['compact', 'full'])
@pytest.mark.parametrize('rho', [0.05, 0.20, 0.50])
def test_reward_config4_subset_vs_oracle(rho, support):
    """BASELINE config 4 as SURVEY 8(d) states it: the 1 M-pair sweep (half of the grids independent draws, half a
    rotated, shifted, 10 % re-randomised copy of the target), of which a seeded 10,000-pair subset is checked
    bit-exact against the oracle -- per density and per support."""
    import gridworld_b200 as gw
    from bench import synth_targets
    n = 1_000_000
    T = synth_targets(n, rho, support, 1237, device='cuda')
    Gd = synth_targets(n, rho, support, 1238, device='cuda')
    h = n // 2
    g = torch.Generator('cuda').manual_seed(99)
    k = int(torch.randint(1, 4, (1,)).item())
    rel = torch.rot90(T[:h], k=k, dims=(2, 3))
    if support == 'compact':   # an admissible translation: the 5 x 5 support stays inside the zone
        rel = torch.roll(rel, shifts=(2, -3), dims=(2, 3))
    noise = torch.rand(rel.shape, device='cuda', generator=g) < 0.1
    rnd = synth_targets(h, 0.3, support, 1239, device='cuda')
    Gd[:h] = torch.where(noise, rnd, rel)
    out, arg = gw.max_intersection(Gd, T, return_argmax=True)
    idx = np.sort(np.random.RandomState(0).choice(n, 10_000, replace=False))
    it = torch.as_tensor(idx, device='cuda')
    Gs, Ts = Gd[it].cpu().numpy(), T[it].cpu().numpy()
    ref, ref_arg = O.max_intersection(Gs, Ts, threads=8, return_argmax=True)
    got, got_arg = out[it].cpu().numpy(), arg[it].cpu().numpy()
    assert np.array_equal(got, ref), np.nonzero(got != ref)[0][:10]
    pos = ref > 0
    assert np.array_equal(got_arg[pos], ref_arg[pos])
    assert (ref[:5000] > ref[5000:].mean()).mean() > 0.5   # the related half really scores higher
    # the sweep without argmax gives the same values
    assert torch.equal(gw.max_intersection(Gd, T), out)


def test_flying_batch_config3_size_vs_oracle():
    """Flying Dict + target_in_obs at config-3 scale per launch (65,536 envs here, the oracle's budget) against the
    oracle: every launch geometry of the big configuration (persistent observe grid, dynamic hand-out) is used."""
    import gridworld_b200 as gw
    n, steps = 65536, 60
    rng = np.random.RandomState(15)
    T, S, poses = _tasks(rng, n)
    env = gw.make_vec(n, action_space='flying', max_steps=40, auto_reset=True, target_in_obs=True,
                      reward_mode='every_step')
    env.set_tasks(T, S, poses)
    orc = O.OracleBatch(n, 'flying', max_steps=40, threads=8)
    orc.set_tasks(T, S, poses)
    bad_envs = np.zeros(n, bool)
    for t in range(steps):
        mv = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
        mv[rng.random_sample((n, 3)) < 0.3] = 0
        cam = rng.uniform(-5, 5, (n, 2)).astype(np.float32)
        inv = rng.randint(0, 7, n).astype(np.int32)
        pl = rng.randint(0, 3, n).astype(np.int32)
        obs, r, d, _ = env.step({'movement': mv, 'camera': cam, 'inventory': inv, 'placement': pl})
        ro, do = orc.step_flying(mv, cam, inv, pl, auto_reset=True)
        if t % 10 == 9 or t == steps - 1:
            s = orc.state()
            pos_err = np.abs(env.position.cpu().numpy() - s['pos']).max(axis=1)
            grid_bad = (env.grid.cpu().numpy() != s['grid']).reshape(n, -1).any(axis=1)
            bad_envs |= (pos_err > POSE_TOL) | grid_bad
        # between state checks an env can only be excused after it was seen diverging
        mism = (r.cpu().numpy() != ro) | (d.cpu().numpy() != do)
        bad_envs |= mism
    assert bad_envs.mean() <= 0.002, bad_envs.sum()
    ok = ~bad_envs
    s = orc.state()
    assert np.array_equal(env.grid.cpu().numpy()[ok], s['grid'][ok])
    assert np.array_equal(env.inventory.cpu().numpy()[ok], s['inventory'][ok])
    assert np.array_equal(env.max_int.cpu().numpy()[ok], s['max_int'][ok])
    assert np.array_equal(obs['target_grid'].cpu().numpy(), T.astype(np.int32))
    assert np.array_equal(obs['grid'].cpu().numpy()[ok], s['grid'][ok].astype(np.int32))


def test_sample_tasks_device_vs_oracle():
    """RandomTasks.sample_task on the device (iglu_sample_tasks) against the oracle's restatement (itself checked
    against the executed reference's distribution): bit-exact structures, shard-splittable stream, and an env whose
    task generator is a device-sampled pool stepping exactly like the oracle on the same pool."""
    import gridworld_b200 as gw
    for spec in (dict(max_blocks=4, height_levels=1, max_dist=2, num_colors=1),
                 dict(max_blocks=6, height_levels=3, max_dist=2, num_colors=4),
                 dict(max_blocks=9, height_levels=2, max_dist=1, num_colors=6),     # blocks that find no free cell
                 dict(max_blocks=1, height_levels=9, max_dist=0, num_colors=2)):
        got = gw.sample_tasks(5000, seed=21, **spec).cpu().numpy()
        want = O.sample_tasks(5000, seed=21, **spec)
        bad = np.argwhere(got != want)
        assert len(bad) == 0, (spec, len(bad), bad[:5].tolist(), got[tuple(bad[0])], want[tuple(bad[0])])
    a = gw.sample_tasks(100, max_blocks=5, height_levels=2, num_colors=3, seed=4, first_task=1000).cpu().numpy()
    b = gw.sample_tasks(1100, max_blocks=5, height_levels=2, num_colors=3, seed=4).cpu().numpy()
    assert np.array_equal(a, b[1000:])
    with pytest.raises(RuntimeError):
        gw.sample_tasks(4, num_colors=7)
    n, m, steps = 1024, 300, 120
    env = gw.make_vec(n, max_steps=20, reward_mode='every_step', target_in_obs=True)
    env.set_random_task_pool(m, max_blocks=5, height_levels=2, max_dist=2, num_colors=3, task_seed=8, seed=3)
    pool = O.sample_tasks(m, max_blocks=5, height_levels=2, max_dist=2, num_colors=3, seed=8)
    orc = O.OracleBatch(n, 'walking', max_steps=20, threads=8)
    orc.set_pool(pool, None, seed=3)
    orc.reset()
    assert np.array_equal(env.task_ids.cpu().numpy(), orc.task_ids)
    rng = np.random.RandomState(2)
    p = np.array([1, 5, 2, 2, 2, 3, 2, 2, 2, 2, 2, 2, 2, 2, 5, 2, 4, 6], float)
    for t in range(steps):
        a = rng.choice(18, n, p=p / p.sum()).astype(np.int32)
        obs, r, d, _ = env.step(a)
        ro, do = orc.step_walking(a, auto_reset=True)
        assert np.array_equal(r.cpu().numpy(), ro) and np.array_equal(d.cpu().numpy(), do), t
    assert np.array_equal(env.grid.cpu().numpy(), orc.state()['grid'])
    assert np.array_equal(obs['target_grid'].cpu().numpy(), orc.target_grids())

    class NoCache:   # the shape of the reference's RandomTasks(max_cache=0): hyper-parameters, no cached tasks
        max_blocks, height_levels, max_dist, num_colors, max_cache, tasks = 3, 1, 2, 2, 0, {}
    env.set_task_generator(NoCache(), seed=6, samples=64)
    assert np.array_equal(env._pool['target'][:, :1089].cpu().numpy().reshape(64, 9, 11, 11),
                          O.sample_tasks(64, max_blocks=3, height_levels=1, max_dist=2, num_colors=2, seed=6))


def test_staged_tasks_golden_on_device():
    """Multi-turn structures with progressive goals on the device (iglu_staged_step) against the reference's own
    Subtasks.step_intersection outputs recorded on an env trace (tests/golden/staged.npz): the episode tasks come from
    the device task pool, the staged right / wrong / done and the goal turn from the staged kernel."""
    import gridworld_b200 as gw
    z = G.load('staged.npz')

    def blocks(dense):   # dense [y+1, x+5, z+5] -> (x, y, z, colour) rows
        return [(int(x) - 5, int(y) - 1, int(zz) - 5, int(dense[y, x, zz])) for y, x, zz in np.argwhere(dense)]
    structures = [[blocks(d) for d in z['seqA']], [blocks(d) for d in z['seqB']]]
    env = gw.make_vec(1, max_steps=1000, auto_reset=False)
    nt = torch.full((1,), -1, dtype=torch.int32, device=env.device)
    env.set_staged_tasks(structures, progressive=True, next_task=nt, reset=False)
    rows = env.staged_rows
    assert rows == [(0, -1), (0, 0), (0, 1), (1, -1)]
    base = env._staged['base'].cpu().numpy()
    last_ep = -1
    for i in range(len(z['ep'])):
        if z['ep'][i] != last_ep:
            nt[0] = rows.index((int(z['structure'][i]), int(z['turn'][i])))
            env.reset()
            last_ep = int(z['ep'][i])
        _, r, _, _ = env.step(z['actions'][i:i + 1])
        sg = env.staged
        got = (int(sg['right'][0]), int(sg['wrong'][0]), int(sg['done'][0]),
               int(sg['goal'][0]) - int(base[int(z['structure'][i])]))
        want = (int(z['right'][i]), int(z['wrong'][i]), int(z['done'][i]), int(z['goal'][i]))
        assert got == want, (i, got, want)
        assert np.array_equal(env.grid[0].reshape(-1).cpu().numpy(), z['grid'][i]), i
        assert float(r[0]) == z['env_reward'][i], i       # the env's own synthetic-task reward is untouched
        assert int(env._staged['state'][2, 0]) == int(z['max_int'][i]), i
    # a batch: every env of a 512-env run agrees with the oracle's Subtasks restatement driven by its own grid
    n = 512
    env = gw.make_vec(n, max_steps=60, auto_reset=True)
    env.set_staged_tasks(structures, progressive=True, seed=5)
    orcs = [None] * n
    rng = np.random.RandomState(8)
    p = np.array([1, 5, 2, 2, 2, 3, 2, 2, 2, 2, 2, 2, 2, 2, 5, 2, 4, 6], float)
    seqs = [z['seqA'], z['seqB']]
    for t in range(150):
        a = rng.choice(18, n, p=p / p.sum()).astype(np.int32)
        env.step(a)
        step_no = env.step_no.cpu().numpy()
        ids = env.task_ids.cpu().numpy()
        grids = env.grid.cpu().numpy().reshape(n, -1)
        sg = {k: v.cpu().numpy() for k, v in env.staged.items()}
        for e in range(0, n, 7):
            if step_no[e] == 1 or orcs[e] is None:
                si, turn = rows[ids[e]]
                orcs[e] = (O.Staged(seqs[si]), si)
                orcs[e][0].reset(turn)
            rr, ww, dd, goal, _ = orcs[e][0].step(grids[e])
            assert (sg['right'][e], sg['wrong'][e], int(sg['done'][e]), sg['goal'][e] - base[orcs[e][1]]) == \
                (rr, ww, dd, goal), (t, e)


def _coded_atlas(rng):
    a = rng.randint(0, 256, (128, 128, 4)).astype(np.uint8)
    a[..., 3] = 255
    return a


@pytest.mark.parametrize('with_atlas,size', [(False, (64, 64)), (True, (64, 64)), (True, (48, 32))])
def test_render_pov_vs_oracle(with_atlas, size):
    """obs['pov'] (csrc/iglu_render.cuh) against its C restatement, bit for bit, on random worlds and poses whose angles
    are multiples of 5 degrees (host-libm trig table); arbitrary angles may flip a pixel on a face edge."""
    import gridworld_b200 as gw
    n = 160
    rng = np.random.RandomState(77)
    atlas = _coded_atlas(rng) if with_atlas else None
    S = _synth(rng, n, 0.08, 'full')
    S[:n // 4] = 0
    T = _synth(rng, n, 0.1, 'compact')
    poses = np.zeros((n, 5))
    poses[:, 0] = rng.randint(-28, 29, n) * 0.25
    poses[:, 1] = rng.randint(-6, 37, n) * 0.25
    poses[:, 2] = rng.randint(-28, 29, n) * 0.25
    poses[:, 3] = rng.randint(0, 72, n) * 5.0
    poses[:, 4] = rng.randint(-18, 19, n) * 5.0
    poses[n // 2:, 3:] += rng.uniform(-2.4, 2.4, (n - n // 2, 2))       # second half: arbitrary angles
    poses[:, 4] = np.clip(poses[:, 4], -90, 90)
    env = gw.make_vec(n, max_steps=50, render=True, render_size=size, texture=atlas, auto_reset=False)
    env.set_tasks(T, S, poses)
    obs = env.reset()
    w, h = size
    assert obs['pov'].shape == (n, h, w, 3) and obs['pov'].dtype == torch.uint8
    pov = obs['pov'].cpu().numpy()
    bad = 0
    for e in range(n):
        want = O.render_pov(S[e], poses[e], h, w, atlas)
        if e < n // 2:
            assert np.array_equal(pov[e], want), e
        else:
            bad += int((pov[e] != want).any(axis=2).sum())
    assert bad <= 1e-3 * (n // 2) * h * w, bad
    assert len(np.unique(pov.reshape(-1, 3), axis=0)) > (20 if with_atlas else 5)     # the frames really show things
    # after a few steps (walking actions keep the angles on the 5 degree lattice for the first half)
    for t in range(6):
        obs, _, _, _ = env.step(rng.randint(0, 18, n).astype(np.int32))
    pov = obs['pov'].cpu().numpy()
    grid = env.grid.cpu().numpy()
    pos, rot = env.position.cpu().numpy(), env.rotation.cpu().numpy()
    for e in range(n // 2):
        want = O.render_pov(grid[e], (*pos[e], *rot[e]), h, w, atlas)
        assert np.array_equal(pov[e], want), e


def test_render_pov_dropin_surface():
    """gym surface with render=True (reference env.py:93-94, 257-258, 299-300): obs['pov'] uint8 (64, 64, 3)."""
    import gridworld_b200 as gw
    env = gw.GridWorld(render=True, vector_state=True, discretize=True, select_and_place=True, max_steps=20, texture=None)
    assert env.observation_space['pov'].shape == (64, 64, 3)
    env.set_task(gw.Task('', gw.Tasks.to_dense([(0, 0, -3, 3)]), starting_grid=[(0, 0, -3, 3), (1, 0, -2, 5)]))
    obs = env.reset()
    assert obs['pov'].shape == (64, 64, 3) and obs['pov'].dtype == np.uint8
    want = O.render_pov(env.grid, (0, 0, 0, 0, 0))
    assert np.array_equal(obs['pov'], want)
    obs, _, _, _ = env.step(13)            # turn the camera: the frame follows
    assert np.array_equal(obs['pov'], O.render_pov(env.grid, (*env.agent.position, *env.agent.rotation)))
    assert not np.array_equal(obs['pov'], want)
    assert np.array_equal(env.render(), obs['pov'])
    fake = gw.GridWorld(render=True, fake=True, vector_state=True, discretize=True)
    assert 'pov' in fake.observation_space.spaces


def test_overlapped_step_counts_every_evaluation_once():
    """The device evaluation counter under the overlapped step (bench.py's evaluated_frac): k_widen evaluates every env
    that holds blocks when the step starts (and is not being reset), k_commit re-evaluates the envs whose block count
    changed and counts only those k_widen did not look at."""
    import gridworld_b200 as gw
    from bench import synth_targets
    n, steps = 1024, 60
    T = synth_targets(n, 0.2, 'compact', 9)
    rng = np.random.RandomState(4)
    env = gw.make_vec(n, max_steps=25, reward_mode='every_step', _sched_bits=OVERLAP)
    env.set_tasks(T)
    env.reset()
    env.reset_evaluation_counter()
    expected = 0
    p = np.array([1, 5, 2, 2, 2, 3, 2, 2, 2, 2, 2, 2, 2, 2, 5, 2, 4, 6], float)
    for t in range(steps):
        ist = env._istate.cpu().numpy()
        dsize_pre, resetting = ist[5].copy(), ((ist[3] >> 16) & 0xFF) != 0
        env.step(rng.choice(18, n, p=p / p.sum()).astype(np.int32))
        ist = env._istate.cpu().numpy()
        dsize_post, changed = ist[5], ist[7] != 0
        expected += int(((dsize_pre != 0) & ~resetting).sum())
        expected += int((changed & (dsize_post != 0) & ((dsize_pre == 0) | resetting)).sum())
    assert env.evaluations == expected and expected > 2000
