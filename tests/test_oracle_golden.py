"""Pins the C oracle (oracle/iglu_oracle.c) against the EXECUTED reference: every golden trace
and reward known answer recorded by oracle/make_golden.py must be reproduced bit-for-bit
(fp64 pose compared with ==, not a tolerance: same libm, same operation order)."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import golden_io as G


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
