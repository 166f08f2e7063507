"""Pins the C oracle (oracle/iglu_oracle.c) against the EXECUTED reference: every golden trace
and reward known answer recorded by oracle/make_golden.py must be reproduced bit-for-bit
(fp64 pose compared with ==, not a tolerance: same libm, same operation order)."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests import golden_io as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _replay(tr, kind):
    b = O.OracleBatch(1, action_space='flying' if kind == 'flying' else 'walking',
                      select_and_place=tr['select_and_place'], max_steps=tr['max_steps'],
                      right_placement_scale=tr['right_scale'], wrong_placement_scale=tr['wrong_scale'])
    b.set_task(0, tr['target'], tr['start'], tr['pose'])
    a = tr['actions']
    for t in range(tr['n']):
        if kind == 'walking':
            r, d = b.step_walking(a['actions'][t:t + 1], auto_reset=True)
        elif kind == 'flying':
            r, d = b.step_flying(a['movement'][t], a['camera'][t], a['inv'][t:t + 1], a['placement'][t:t + 1],
                                 auto_reset=True)
        else:
            r, d = b.step_walking_dict(a['buttons'][t], a['camera'][t], auto_reset=True)
        s = b.state()
        ctx = f"{tr['name']} step {t}"
        assert s['pos'][0].tolist() == tr['pos'][t].tolist(), ctx
        assert s['rot'][0].tolist() == tr['rot'][t].tolist(), ctx
        assert s['dy'][0] == tr['dy'][t], ctx
        assert s['tis'][0] == tr['tis'][t], ctx
        assert s['active'][0] == tr['active'][t], ctx
        assert s['inventory'][0].tolist() == tr['inventory'][t].tolist(), ctx
        assert r[0] == tr['reward'][t], ctx
        assert bool(d[0]) == bool(tr['done'][t]), ctx
        assert s['max_int'][0] == tr['max_int'][t], ctx
        assert s['step_no'][0] == tr['step_no'][t], ctx
        assert np.array_equal(s['grid'][0].reshape(-1), G.grid_at(tr, t)), ctx
        if t % 16 == 0 or t == tr['n'] - 1:
            o = b.obs()
            assert np.array_equal(o['agentPos'][0], tr['agentPos'][t]), ctx
            assert np.array_equal(o['compass'][0], tr['compass'][t]), ctx
            assert np.array_equal(o['inventory'][0], tr['obs_inventory'][t]), ctx


@pytest.mark.parametrize('tr', G.walking_traces(), ids=lambda t: t['name'])
def test_walking_traces(tr):
    _replay(tr, 'walking')


@pytest.mark.parametrize('tr', G.flying_traces(), ids=lambda t: t['name'])
def test_flying_traces(tr):
    _replay(tr, 'flying')


@pytest.mark.parametrize('tr', G.dict_traces(), ids=lambda t: t['name'])
def test_walking_dict_traces(tr):
    _replay(tr, 'dict')


@pytest.mark.parametrize('tr', G.generator_traces(), ids=lambda t: t['name'])
def test_generator_traces(tr):
    """env.set_task_generator: every reset draws the next task first (reference env.py:213-216). The indices the
    reference's generator drew (numpy global stream) are imposed through next_task."""
    b = O.OracleBatch(1, action_space=tr['space'], max_steps=tr['max_steps'])
    b.set_pool(tr['pool_targets'], tr['pool_starts'], seed=3)
    ch = list(tr['choices'])
    b.next_task[0] = ch.pop(0)
    b.reset()
    a = tr['actions']
    done = False
    for t in range(tr['n']):
        if done:
            b.next_task[0] = ch.pop(0)
        if tr['space'] == 'walking':
            r, d = b.step_walking(a['actions'][t:t + 1], auto_reset=True)
        else:
            r, d = b.step_flying(a['movement'][t], a['camera'][t], a['inv'][t:t + 1], a['placement'][t:t + 1],
                                 auto_reset=True)
        done = bool(d[0])
        s = b.state()
        ctx = f"{tr['name']} step {t}"
        assert b.task_ids[0] == tr['task_at_step'][t], ctx
        assert s['pos'][0].tolist() == tr['pos'][t].tolist(), ctx
        assert s['rot'][0].tolist() == tr['rot'][t].tolist(), ctx
        assert s['dy'][0] == tr['dy'][t], ctx
        assert s['inventory'][0].tolist() == tr['inventory'][t].tolist(), ctx
        assert r[0] == tr['reward'][t], ctx
        assert done == bool(tr['done'][t]), ctx
        assert s['max_int'][0] == tr['max_int'][t], ctx
        assert s['step_no'][0] == tr['step_no'][t], ctx
        assert np.array_equal(s['grid'][0].reshape(-1), G.grid_at(tr, t)), ctx
        assert np.array_equal(b.target_grids()[0], tr['pool_targets'][tr['task_at_step'][t]]), ctx
    assert not ch


def test_generator_draw_formula():
    """The counter-based draw published in include/iglu_b200.h, restated in plain Python integers."""
    M64 = (1 << 64) - 1

    def draw(seed, env, ep, m):
        x = (seed + 0x9E3779B97F4A7C15 * (env + 1) + 0xD1B54A32D192ED03 * ep) & M64
        x ^= x >> 30
        x = (x * 0xBF58476D1CE4E5B9) & M64
        x ^= x >> 27
        x = (x * 0x94D049BB133111EB) & M64
        x ^= x >> 31
        return ((x >> 32) * m) >> 32

    n, m, seed, off = 64, 13, 0xDEADBEEFCAFE1234, 1000
    rng = np.random.RandomState(0)
    T = (rng.rand(m, 9, 11, 11) < 0.05).astype(np.int8)
    b = O.OracleBatch(n, max_steps=2)
    b.set_pool(T, None, seed=seed, env_offset=off)
    b.reset()
    ids = lambda ep: [draw(seed, off + e, ep, m) for e in range(n)]  # noqa: E731
    assert b.task_ids.tolist() == ids(0)
    for _ in range(3):  # max_steps = 2: done on the 2nd step, auto-reset (= next draw) inside the 3rd
        b.step_walking(np.zeros(n, np.int32), auto_reset=True)
    assert b.task_ids.tolist() == ids(1)
    b.reset()  # an explicit reset draws as well; the episode counter keeps counting
    assert b.task_ids.tolist() == ids(2)
    b.next_task[5] = 7  # imposed index: used once, then consumed
    b.reset()
    want = ids(3)
    want[5] = 7
    assert b.task_ids.tolist() == want and b.next_task[5] == -1
    assert len(set(ids(0) + ids(1) + ids(2))) > m // 2


def test_goals_known_answers():
    z = G.load('goals.npz')
    T = z['targets']
    n = len(T)
    g1 = np.stack([np.rot90(t, k=-1, axes=(1, 2)) for t in T])
    assert np.array_equal(O.max_intersection(g1, T), z['ka1'])
    # SURVEY section 4 table (C1, C3, C8, C17, C32, C60, C96, C118)
    names = list(z['names'])
    for nm, blocks, ka1, ka2 in (('C1', 29, 29, 15), ('C3', 3, 3, 2), ('C8', 12, 12, 6), ('C17', 3, 3, 2),
                                 ('C32', 5, 5, 3), ('C60', 20, 20, 0), ('C96', 16, 16, 2), ('C118', 12, 12, 6)):
        i = names.index(nm)
        assert z['target_size'][i] == blocks and z['ka1'][i] == ka1
    # KA2 grids (every 2nd block shifted +1 in x), rebuilt exactly as oracle/make_golden.py does
    g2 = np.zeros_like(T)
    for i, t in enumerate(T):
        for k, (y, x, zz) in enumerate(np.argwhere(t)):
            if k % 2 == 1:
                if x + 1 <= 10:
                    g2[i, y, x + 1, zz] = t[y, x, zz]
            else:
                g2[i, y, x, zz] = t[y, x, zz]
    assert np.array_equal(O.max_intersection(g2, T), z['ka2'])
    for j in range(z['grids'].shape[1]):
        mi, arg = O.max_intersection(z['grids'][:, j], T, return_argmax=True)
        assert np.array_equal(mi, z['max_int'][:, j])
        assert np.array_equal(arg, z['argmax'][:, j])
    for i in range(n):
        n_adm, rect, ts = O.task_describe(T[i])
        assert n_adm.tolist() == z['n_adm'][i].tolist()
        assert ts == z['target_size'][i]
        # the admissible set is the full rectangle
        for r in range(4):
            assert (rect[r, 1] - rect[r, 0] + 1) * (rect[r, 3] - rect[r, 2] + 1) == n_adm[r]


def test_reward_pairs():
    z = G.load('reward_pairs.npz')
    for i in range(len(z['max_int'])):
        full = z['fulls'][i:i + 1] if z['has_full'][i] else None
        mi, arg = O.max_intersection(z['grids'][i:i + 1], z['targets'][i:i + 1], full, bool(z['invariant'][i]),
                                     return_argmax=True)
        assert mi[0] == z['max_int'][i], i
        assert arg[0].tolist() == z['argmax'][i].tolist(), i
        n_adm, rect, ts = O.task_describe(z['targets'][i], None if full is None else full[0], bool(z['invariant'][i]))
        assert n_adm.tolist() == z['n_adm'][i].tolist(), i
        assert ts == z['target_size'][i]


@pytest.mark.skipif(not os.path.isdir('/root/reference/gridworld'), reason='needs the reference tree (build container)')
def test_sample_tasks_matches_reference_distribution():
    """The oracle's sample_task restatement (counter-based stream) against the executed RandomTasks.sample_task
    (task_set.py:135-157, numpy's Mersenne Twister): same structure invariants, same distributions."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, 'oracle', 'refharness'))
    import load_reference
    load_reference.load()
    from gridworld.tasks.task_set import RandomTasks
    n = 1500
    for spec in (dict(max_blocks=4, height_levels=1, max_dist=2, num_colors=1),
                 dict(max_blocks=6, height_levels=3, max_dist=2, num_colors=4)):
        np.random.seed(11)
        gen = RandomTasks(max_cache=0, **spec)
        ref = np.stack([np.asarray(gen.sample_task().target_grid) for _ in range(n)]).astype(np.int8)
        got = O.sample_tasks(n, seed=5, **spec)

        def stats(T):
            lv = T[:, :spec['height_levels']]
            assert (T[:, spec['height_levels']:] == 0).all()
            cnt = (lv != 0).reshape(len(T), spec['height_levels'], -1).sum(2)
            assert (cnt == spec['max_blocks']).all()          # every level gets exactly max_blocks blocks
            assert T.max() <= spec['num_colors'] and T.min() >= 0
            ext = []
            for t in lv.reshape(-1, 11, 11):
                xs, zs = np.nonzero(t)
                ext.append((xs.max() - xs.min(), zs.max() - zs.min()))
            ext = np.array(ext)
            assert ext.max() <= 2 * spec['max_dist']           # all blocks within max_dist of the level's first block
            occ = (lv != 0).reshape(-1, 11, 11).mean(0)         # where blocks land
            col = np.bincount(lv[lv != 0].ravel(), minlength=7)[1:] / np.count_nonzero(lv)
            hx = np.bincount(ext[:, 0], minlength=5) / len(ext)
            hz = np.bincount(ext[:, 1], minlength=5) / len(ext)
            return occ, col, hx, hz
        for a, b in zip(stats(ref), stats(got)):
            assert np.abs(a - b).max() < 0.05, (spec, np.abs(a - b).max())
    # the stream is a function of (seed, task index): shards can split a pool
    whole = O.sample_tasks(64, max_blocks=5, height_levels=2, num_colors=3, seed=9)
    part = O.sample_tasks(24, max_blocks=5, height_levels=2, num_colors=3, seed=9, first_task=40)
    assert np.array_equal(whole[40:], part)


def test_staged_tasks_golden():
    """Subtasks.step_intersection with progressive goals (task.py:288-298) as recorded from the reference on the
    grids of an env trace (tests/golden/staged.npz) vs the oracle's restatement."""
    z = G.load('staged.npz')
    objs = [O.Staged(z['seqA']), O.Staged(z['seqB'])]
    last_ep = -1
    for i in range(len(z['ep'])):
        st = objs[int(z['structure'][i])]
        if z['ep'][i] != last_ep:
            st.reset(int(z['turn'][i]))
            last_ep = int(z['ep'][i])
        got = st.step(z['grid'][i])
        want = (int(z['right'][i]), int(z['wrong'][i]), int(z['done'][i]), int(z['goal'][i]), int(z['max_int'][i]))
        assert got == want, (i, got, want)
    assert set(z['goal'].tolist()) == {0, 1, 2, 3}   # every turn incl. the one past the last was reached


def _coded_atlas():
    """A 128 x 128 RGBA atlas whose texel (tx, ty) -- ty counted from the BOTTOM like OpenGL -- holds (tx, ty, 7)."""
    a = np.zeros((128, 128, 4), np.uint8)
    ty, tx = np.meshgrid(np.arange(128), np.arange(128), indexing='ij')
    a[127 - ty, tx, 0] = tx
    a[127 - ty, tx, 1] = ty
    a[..., 2] = 7
    a[..., 3] = 255
    return a


def _box_mask(block, pose, h=64, w=64):
    """Pixels whose centre ray hits the unit cube centred on `block` in front of the near plane (numpy slab test)."""
    x, y, z, yaw, pitch = pose
    sy_, cy = np.sin(np.radians(yaw)), np.cos(np.radians(yaw))
    sp, cp = np.sin(np.radians(pitch)), np.cos(np.radians(pitch))
    f, r = np.array([sy_ * cp, sp, -cy * cp]), np.array([cy, 0, sy_])
    u = np.cross(r, f)
    i, j = (np.arange(w) + 0.5) * 2 / w - 1, 1 - (np.arange(h) + 0.5) * 2 / h
    d = f[None, None, :] + i[None, :, None] * r[None, None, :] * (w / h) + j[:, None, None] * u[None, None, :]
    o = np.array([x, y, z], float)
    lo = np.array(block, float) - 0.5
    with np.errstate(divide='ignore', invalid='ignore'):
        t1, t2 = (lo - o) / d, (lo + 1 - o) / d
    tn, tf = np.minimum(t1, t2).max(axis=2), np.maximum(t1, t2).min(axis=2)
    return (tn <= tf) & (tn >= 0.1) & (tn <= 30)


def test_render_pov_known_answers():
    """obs['pov'] restatement (reference gridworld/render.py:94-113, 159-170; utils.py:26-41, 75-139): answers that
    follow from the camera / geometry / texture rules by hand."""
    SKY, WHITE_, GREY_, RED_ = (128, 176, 255), (245, 245, 245), (70, 72, 76), (133, 16, 17)
    empty = np.zeros((9, 11, 11), np.int8)
    # straight up: only sky; straight down from the origin: only the white build-zone floor (1.5 below, 90 degree fov)
    assert (O.render_pov(empty, (0, 0, 0, 0, 90)) == SKY).all()
    assert (O.render_pov(empty, (0, 0, 0, 35, -90)) == WHITE_).all()
    # level view: sky above the horizon and just under it (row 33: the floor lies beyond the far plane, 1.5 / (1.5 / 32)
    # = 32 > 30; row 34: beyond the floor's edge, 19.2 > 18.5), then grey floor, white build-zone floor near the bottom
    img = O.render_pov(empty, (0, 0, 0, 0, 0))
    assert (img[:35] == SKY).all() and (img[35] == GREY_).all() and (img[63] == WHITE_).all()
    # below the floor nothing is drawn (top faces are culled from underneath)
    assert (O.render_pov(empty, (0, -3, 0, 0, -45)) == SKY).all()
    # one red block at (0, -1, -3), i.e. its front face 2.5 ahead of an eye at the origin looking along -z: the face
    # spans sx in [-0.2, 0.2] -> columns 26..37 and sy in [-0.6, -0.2] -> rows 38..50; its top face (y = -0.5, seen from
    # above) fills row 37 (sy = -0.172: the ray reaches y = -0.5 at depth 2.91, where the block is 0.344 wide in sx ->
    # columns 27..36, plus the two rays that graze the block's vertical edges exactly: ties, either way is right)
    g = empty.copy()
    g[0, 5, 2] = 3
    img = O.render_pov(g, (0, 0, 0, 0, 0))
    red = (img == RED_).all(axis=2)
    ys, xs = np.nonzero(red)
    assert (xs.min(), xs.max(), ys.min(), ys.max()) == (26, 37, 37, 50), (xs.min(), xs.max(), ys.min(), ys.max())
    assert red[37, 27:37].all() and red[38:51, 26:38].all() and red.sum() in (166, 168)
    assert (red[:, :32] == red[:, :31:-1]).all()              # mirror symmetry
    # the same picture from the slab test of the block's box, for a few poses (independent of the cell walk)
    for pose in ((0, 0, 0, 0, 0), (0.3, 1.2, 1.0, 20, -25), (-1.7, 2.4, 0.9, 340, -35), (2, 0.5, -1, 215, -10)):
        got = (O.render_pov(g, pose) == RED_).all(axis=2)
        assert (got != _box_mask((0, -1, -3), pose)).sum() <= 2, pose
    # ... also when other blocks enlarge the walked region (far corners, not in front of the red block)
    g_far = g.copy()
    g_far[8, 0, 10] = 1
    g_far[8, 10, 10] = 1
    got = (O.render_pov(g_far, (0.3, 1.2, 1.0, 20, -25)) == RED_).all(axis=2)
    assert (got != _box_mask((0, -1, -3), (0.3, 1.2, 1.0, 20, -25))).sum() <= 2
    # the same block seen after turning: yaw 90 looks along +x, so the block has to be at x = +3 to show the same picture
    g2 = empty.copy()
    g2[0, 8, 5] = 3
    # (cos 90 degrees is 6e-17, not 0, in fp64: a couple of pixels on face edges may flip)
    assert (O.render_pov(g2, (0, 0, 0, 90, 0)) != img).any(axis=2).sum() <= 4
    # inside a block (or closer than the near plane to its face) the block is not seen
    g3 = empty.copy()
    g3[1, 5, 5] = 3                                           # the cell the eye is in
    assert np.array_equal(O.render_pov(g3, (0, 0, 0, 0, 0)), O.render_pov(empty, (0, 0, 0, 0, 0)))
    # texel addressing: white floor tile (0, 0) uses the whole 32 x 32 tile, u with z and v with x
    atlas = _coded_atlas()
    img = O.render_pov(empty, (0.25, 0, -0.25, 0, -90), atlas=atlas)
    c = img[32, 32]                                           # ray through (0.25 + eps, -1.5, -0.25 + eps') of cell (0, 0)
    assert c[2] == 7 and c[0] in (7, 8) and c[1] in (23, 24), c     # u = z + .5 = .25 -> texel 8; v = x + .5 = .75 -> 24
    # a colour block's side faces use quarter (0, 0) of its tile, the top quarter (0, 16): RED is tile (0, 1)
    img = O.render_pov(g, (0, -1, 0, 0, 0), atlas=atlas)      # eye level with the block centre: its front (+z) face
    c = img[32, 32]
    assert 16 <= c[0] < 32 and 32 + 16 <= c[1] < 64, c        # front / back faces: quarter (16, 16) of tile (0, 1)
    img = O.render_pov(g, (0, 3, -3, 0, -90), atlas=atlas)    # from above: top face, quarter (0, 16)
    c = img[32, 32]
    assert 0 <= c[0] < 16 and 32 + 16 <= c[1] < 64, c
