"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol of include/iglu_b200.h,
the host layer fails loudly without a CUDA device, host task objects follow the reference's
Task.__init__ products, ctypes struct layouts match the header."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tests import golden_io as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from gridworld_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, 'include', 'iglu_b200.h')).read()
    declared = set(re.findall(r'\b(iglu_[a-z_]+)\s*\(', header))
    declared -= {'iglu_ctx'}
    assert declared == set(_lib.EXPORTS)
    for sym in declared:
        assert getattr(lib, sym) is not None
    assert lib.iglu_abi_version() == 1
    assert lib.iglu_strerror(0) == b'ok'
    assert b'Task is not initialized' in lib.iglu_strerror(-3)


def test_struct_layouts_match_header():
    from gridworld_b200 import _lib
    assert C.sizeof(_lib.Config) == 8 * 4 + 2 * 8
    assert C.sizeof(_lib.State) == 8 * 11
    assert C.sizeof(_lib.TaskPool) == 8 * 9
    assert C.sizeof(_lib.Obs) == 8 * 7
    assert _lib.GRID_STRIDE % 16 == 0 and _lib.GRID_STRIDE >= _lib.CELLS


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    import gridworld_b200 as gw
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        gw.make_vec(4)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        gw.make('IGLUGridworld-v0', render=False)
    with pytest.raises(RuntimeError):
        gw.max_intersection(np.zeros((1, 9, 11, 11)), np.zeros((1, 9, 11, 11)))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        gw.GridWorld(render=True)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'gridworld_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, re.M), f
                assert 'libiglu_oracle' not in src and 'iglu_oracle' not in src, f


def test_host_task_matches_reference_products():
    """Task.admissible / target_grids (closed forms) against the reference's Task.__init__ outputs."""
    import gridworld_b200 as gw
    z = G.load('goals.npz')
    for i in range(0, len(z['targets']), 7):
        t = gw.Task('', z['targets'][i].astype(np.int32))
        assert [len(a) for a in t.admissible] == z['n_adm'][i].tolist()
        assert t.target_size == z['target_size'][i]
    zp = G.load('reward_pairs.npz')
    for i in range(len(zp['max_int'])):
        full = zp['fulls'][i].astype(np.int32) if zp['has_full'][i] else None
        t = gw.Task('', zp['targets'][i].astype(np.int32), full_grid=full, invariant=bool(zp['invariant'][i]))
        n_adm = [len(a) for a in t.admissible]
        assert n_adm + [0] * (4 - len(n_adm)) == zp['n_adm'][i].tolist()
    t = gw.Task('', z['targets'][0].astype(np.int32))
    r1 = t.target_grids[1]
    for x in range(11):
        for zz in range(11):
            assert np.array_equal(r1[:, zz, 10 - x], z['targets'][0][:, x, zz])  # task.py:52-53


def test_to_dense_to_sparse():
    import gridworld_b200 as gw
    d = gw.Tasks.to_dense([(0, -1, -2, 1), (5, 7, 5, 6)])
    assert d.shape == (9, 11, 11) and d[0, 5, 3] == 1 and d[8, 10, 10] == 6
    sp = gw.Tasks.to_sparse(np.asarray(d))
    assert sorted(sp) == sorted([(0 - 5, 5 - 1, 3 - 5, 1), (8 - 5, 10 - 1, 10 - 5, 6)])  # axis-0-as-x quirk
    # a cell listed twice keeps the last colour; non-list inputs pass through; to_sparse leaves lists alone
    assert gw.Tasks.to_dense([(1, 0, 1, 2), (1, 0, 1, 5)])[1, 6, 6] == 5
    assert gw.Tasks.to_dense(d) is d and gw.Tasks.to_dense([]).sum() == 0
    assert gw.Tasks.to_sparse([(1, 2, 3, 4)]) == [(1, 2, 3, 4)]


@pytest.mark.skipif(not os.path.isdir('/root/reference/gridworld'), reason='needs the reference tree (build container)')
def test_converters_and_collections_match_live_reference():
    """Tasks.to_dense / to_sparse against the executed reference, and the reference's OWN collection objects
    (CustomTasks, RandomTasks, Subtasks, Task) fed through enumerate_tasks / pool_arrays by duck-typing."""
    import sys
    import gridworld_b200 as gw
    sys.path.insert(0, os.path.join(ROOT, 'oracle', 'refharness'))
    import load_reference
    load_reference.load()
    from gridworld.tasks import task as rtask, task_set as rset
    rng = np.random.RandomState(3)
    for _ in range(20):
        k = rng.randint(0, 30)
        blocks = [(int(rng.randint(-5, 6)), int(rng.randint(-1, 8)), int(rng.randint(-5, 6)), int(rng.randint(1, 7)))
                  for _ in range(k)]
        ref_d = rtask.Tasks.to_dense(blocks)
        assert np.array_equal(gw.Tasks.to_dense(blocks), ref_d)
        assert [tuple(int(v) for v in r) for r in rtask.Tasks.to_sparse(np.asarray(ref_d))] == gw.Tasks.to_sparse(np.asarray(ref_d))
    np.random.seed(0)
    goals = [('a', [(0, -1, 0, 1), (1, -1, 0, 2)]), ('b', [(2, 0, 2, 3)])]
    ct = rset.CustomTasks(goals)
    T, S, chats = gw.pool_arrays(ct)
    assert T.shape == (2, 9, 11, 11) and sorted(chats) == ['a', 'b'] and S.sum() == 0
    assert sorted(int(t.sum()) for t in T) == [3, 3]
    rt = rset.RandomTasks(max_blocks=4, height_levels=2, num_colors=3, max_cache=5)
    T, S, _ = gw.pool_arrays(rt)
    assert T.shape[0] == 5 and all(np.count_nonzero(t) == 8 for t in T)
    rt0 = rset.RandomTasks(max_blocks=3, height_levels=1, max_cache=0)
    assert gw.pool_arrays(rt0, samples=7)[0].shape[0] == 7
    seq = [[(0, -1, 0, 1)], [(0, -1, 0, 1), (1, -1, 0, 2)], [(0, -1, 0, 1), (1, -1, 0, 2), (1, 0, 0, 3)]]
    sub = rtask.Subtasks(['x', 'y', 'z'], seq)
    T, S, chats = gw.pool_arrays(sub)
    assert [np.count_nonzero(t) for t in T] == [1, 2, 3] and [np.count_nonzero(s) for s in S] == [0, 1, 2]
    one = rtask.Task('hi', ref_d, starting_grid=[(0, 0, 0, 1)])
    T, S, chats = gw.pool_arrays(one)
    assert chats == ['hi'] and S[0, 1, 5, 5] == 1 and np.array_equal(T[0], ref_d.astype(np.int8))


def test_spaces():
    from gridworld_b200.vec_env import make_action_space, make_observation_space
    assert make_action_space('walking', True).n == 18
    fl = make_action_space('flying', True)
    assert set(fl.spaces) == {'movement', 'camera', 'inventory', 'placement'}
    wd = make_action_space('walking', False)
    assert set(wd.spaces) == {'forward', 'back', 'left', 'right', 'jump', 'attack', 'use', 'camera', 'hotbar'}
    ob = make_observation_space(True, True)
    assert ob['grid'].shape == (9, 11, 11) and ob['grid'].dtype == np.int32
    assert ob['agentPos'].shape == (5,) and ob['target_grid'].shape == (9, 11, 11)
    assert fl.contains(fl.sample())


def test_pool_arrays_and_goal_file_loader(tmp_path):
    """Task collections -> the arrays IGLUVecEnv.set_task_pool takes; goals.pkl-format loader (restricted unpickler)."""
    import pickle
    from gridworld_b200.tasks import Task, Tasks, enumerate_tasks, load_goals, pool_arrays
    from tests import golden_io as G
    z = G.load('goals.npz')
    dense = z['targets'][:12]
    goals = {}
    for nm, t in zip(z['names'][:12], dense):
        goals[str(nm)] = [(np.int64(x - 5), np.int64(y - 1), np.int64(zz - 5), np.int64(t[y, x, zz]))
                          for y, x, zz in np.argwhere(t)]
    path = tmp_path / 'goals.pkl'
    path.write_bytes(pickle.dumps(goals))
    loaded = load_goals(str(path))
    assert [nm for nm, _ in loaded] == sorted(goals, key=lambda s: int(s[1:]))
    for nm, blocks in loaded:
        assert all(type(v) is int for b in blocks for v in b)
        assert np.array_equal(np.asarray(Tasks.to_dense(blocks)).astype(np.int8), dense[list(z['names'][:12]).index(nm)])
    # pool arrays of a plain list of Tasks (with and without starting grids) and of a CustomTasks collection
    start = [(0, -1, 1, 3), (1, -1, 1, 4)]
    tasks = [Task(nm, np.asarray(Tasks.to_dense(blocks)), starting_grid=start if i % 2 else [])
             for i, (nm, blocks) in enumerate(loaded)]
    T, S, chats = pool_arrays(tasks)
    order = [list(z['names'][:12]).index(nm) for nm, _ in loaded]
    assert T.dtype == np.int8 and np.array_equal(T, dense[order]) and chats == [nm for nm, _ in loaded]
    assert S[0].sum() == 0 and S[1][0, 5, 6] == 3 and S[1][0, 6, 6] == 4 and np.count_nonzero(S[1]) == 2
    class Bag:  # any object with a `.tasks` dict is a collection (the shape of the reference's CustomTasks)
        def __init__(self, ts):
            self.tasks = {str(i): t for i, t in enumerate(ts)}
    assert len(enumerate_tasks(Bag(tasks))) == 12
    assert np.array_equal(pool_arrays(Bag(tasks))[0], T)
    # dense starting grids are accepted as they are
    dense_start = np.zeros((9, 11, 11), int)
    dense_start[0, 5, 6] = 3
    assert pool_arrays([Task('', dense[0], starting_grid=dense_start)])[1][0, 0, 5, 6] == 3
    # a pickle that wants to import anything else is refused before it can run
    class Evil:
        def __reduce__(self):
            import os
            return (os.system, ('echo pwned',))
    bad = tmp_path / 'bad.pkl'
    bad.write_bytes(pickle.dumps({'C1': Evil()}))
    with pytest.raises(pickle.UnpicklingError):
        load_goals(str(bad))


def test_tasks_dataset_loader(tmp_path):
    """IGLUDataset-format parsed-task cache (bz2-pickled {task id: [Subtasks]}, reference data/iglu_dataset.py:204-216)
    read without the reference: the single-turn tasks equal the ones the reference's Subtasks.__iter__ produced."""
    import bz2
    import pickle
    from gridworld_b200.tasks import Tasks, load_tasks_dataset, pool_arrays
    z = G.load('tasks_dataset_expected.npz')
    loaded = load_tasks_dataset(os.path.join(G.GOLDEN, 'tasks_dataset.pkl.bz2'))
    tasks = [(nm, k, t) for nm, k, ts in loaded for t in ts]
    assert [t[0] for t in tasks] == z['names'].tolist() and [t[1] for t in tasks] == z['session'].tolist()
    for i, (_, _, t) in enumerate(tasks):
        assert np.array_equal(t.target_grid, z['targets'][i]) and t.chat == str(z['chats'][i])
        assert np.array_equal(np.asarray(Tasks.to_dense(t.starting_grid)), z['starts'][i])
        assert np.array_equal(t.full_grid, z['fulls'][i])
    T, S, chats = pool_arrays([t for _, _, t in tasks])
    assert np.array_equal(T, z['targets']) and np.array_equal(S, z['starts']) and chats == z['chats'].tolist()

    class Evil:
        def __reduce__(self):
            import os as _os
            return (_os.system, ('echo pwned',))
    bad = tmp_path / 'bad.pkl.bz2'
    bad.write_bytes(bz2.compress(pickle.dumps({'c1': [Evil()]})))
    with pytest.raises(pickle.UnpicklingError):
        load_tasks_dataset(str(bad))


def test_staged_arrays_follow_subtasks_enumeration():
    """tasks.staged_arrays (the inputs of IGLUVecEnv.set_staged_tasks): episode tasks = what Subtasks.reset() can
    create, stage rows = the per-turn targets + one repeat of the last, all scored against the full structure."""
    from gridworld_b200.tasks import Tasks, staged_arrays
    z = G.load('staged.npz')

    def blocks(dense):
        return [(int(x) - 5, int(y) - 1, int(zz) - 5, int(dense[y, x, zz])) for y, x, zz in np.argwhere(dense)]
    A, B = [blocks(d) for d in z['seqA']], [blocks(d) for d in z['seqB']]

    class Sub:   # the shape of the reference's Subtasks
        def __init__(self, seq, dialog):
            self.structure_seq, self.dialog = seq, dialog
    a = staged_arrays([Sub(A, ['one', ['two', 'ok'], 'three']), B])
    assert a['rows'] == [(0, -1), (0, 0), (0, 1), (1, -1)]
    assert np.array_equal(a['pool_targets'], np.stack([z['seqA'][0], z['seqA'][1], z['seqA'][2], z['seqB'][0]]))
    assert a['pool_starts'][0].sum() == 0 and np.array_equal(a['pool_starts'][1], z['seqA'][0])
    assert np.array_equal(a['pool_starts'][2], z['seqA'][1]) and a['pool_starts'][3].sum() == 0
    assert a['chats'] == ['one', 'one\ntwo\nok', 'one\ntwo\nok\nthree', '']
    assert a['next_stage'].tolist() == [1, 2, 3, -1, 5, -1] and a['stage_of'].tolist() == [0, 1, 2, 4]
    assert np.array_equal(a['stage_targets'][3], z['seqA'][2]) and np.array_equal(a['stage_targets'][5], z['seqB'][0])
    assert all(np.array_equal(f, z['seqA'][2]) for f in a['stage_fulls'][:4])
    assert a['stage_base'].tolist() == [0, 4]
    assert np.array_equal(Tasks.to_dense(A[2]), z['seqA'][2])


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the GPU arm) prints ONE JSON line with the
    contract's keys; it never touches CUDA."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '3',
                          '--warmup', '1'], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'env_steps_per_sec' and d['unit'] == 'env-steps/s'
    assert d['higher_is_better'] is True and d['steps'] == 3 and d['value'] > 0 and d['ms_per_step'] > 0
    assert d['cpu_baseline']['kind'] in ('port', 'reference') and d['cpu_baseline']['cores'] >= 1
    assert d['cpu_port']['kind'] == 'port' and d['cpu_port']['value'] > 0
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config']


def test_load_texture_locates_an_installed_reference(tmp_path, monkeypatch):
    """load_texture('auto') finds gridworld/texture.png of an installed reference package without importing it, and
    answers None when there is none."""
    import sys
    import gridworld_b200 as gw
    pytest.importorskip('PIL')
    from PIL import Image
    monkeypatch.setattr(sys, 'path', [p for p in sys.path if not os.path.isdir(os.path.join(p, 'gridworld'))])
    sys.modules.pop('gridworld', None)
    assert gw.load_texture('auto') is None
    pkg = tmp_path / 'gridworld'
    pkg.mkdir()
    (pkg / '__init__.py').write_text('raise RuntimeError("must not be imported")\n')
    rng = np.random.RandomState(0)
    atlas = rng.randint(0, 256, (128, 128, 4)).astype(np.uint8)
    Image.fromarray(atlas, 'RGBA').save(pkg / 'texture.png')
    monkeypatch.syspath_prepend(str(tmp_path))
    got = gw.load_texture('auto')
    assert got is not None and got.shape == (128, 128, 4) and np.array_equal(got, atlas)
    assert np.array_equal(gw.load_texture(str(pkg / 'texture.png')), atlas)
