"""Host-side task objects with the reference's interface (reference gridworld/tasks/task.py,
task_set.py). They only carry data; maximal_intersection / argmax_intersection run on the GPU
through iglu_max_intersection, and inside an env the synthetic task is prepared on the device by
iglu_set_tasks (rotations, admissible rectangles, block list)."""
import io
import pickle

import numpy as np

BUILD_ZONE_SIZE_X = 11
BUILD_ZONE_SIZE_Z = 11
BUILD_ZONE_SIZE = 9, 11, 11


def _bbox_rects(grid, n_rot=4):
    """Admissible shift rectangles (dx_lo, dx_hi, dz_lo, dz_hi) per rotation from the bounding box of
    the non-zero cells -- the closed form of reference task.py:57-72."""
    nz = np.argwhere(np.asarray(grid) != 0)
    if len(nz) == 0:
        xmin, xmax, zmin, zmax = 10, 0, 10, 0
    else:
        xmin, xmax = int(nz[:, 1].min()), int(nz[:, 1].max())
        zmin, zmax = int(nz[:, 2].min()), int(nz[:, 2].max())
    rects = []
    for _ in range(n_rot):
        rects.append((xmax - 10, xmin, zmax - 10, zmin))
        xmin, xmax, zmin, zmax = zmin, zmax, 10 - xmax, 10 - xmin
    return rects


class Task:
    """reference gridworld/tasks/task.py:8-161"""

    def __init__(self, chat, target_grid, last_instruction=None, starting_grid=None, full_grid=None,
                 invariant=True):
        self.chat = chat
        self.starting_grid = starting_grid
        self.last_instruction = last_instruction
        self.full_grid = full_grid
        self.invariant = invariant
        target_grid = np.asarray(target_grid)
        self.target_size = int((target_grid != 0).sum())
        self.full_size = self.target_size
        if full_grid is not None:
            self.full_size = int((np.asarray(full_grid) != 0).sum())
        self.target_grid = target_grid
        self.max_int = 0
        self.prev_grid_size = 0
        self.right_placement = 0
        self.wrong_placement = 0

    # -- products of the reference's __init__, computed lazily (not needed on the device path) --
    @property
    def target_grids(self):
        """four y-rotations, T_k[:, z, 10 - x] = T_{k-1}[:, x, z] (task.py:47-56)"""
        return [np.rot90(self.target_grid, k=-k, axes=(1, 2)).copy() for k in range(4)]

    @property
    def admissible(self):
        if not self.invariant:
            return [[(0, 0)]]
        base = self.full_grid if self.full_grid is not None else self.target_grid
        return [[(dx, dz) for dx in range(r[0], r[1] + 1) for dz in range(r[2], r[3] + 1)]
                for r in _bbox_rects(base)]

    def reset(self):
        """task.py:74-86"""
        if self.starting_grid is not None:
            self.max_int = self.maximal_intersection(Tasks.to_dense(self.starting_grid))
        else:
            self.max_int = 0
        self.prev_grid_size = len(self.starting_grid) if self.starting_grid is not None else 0
        self.right_placement = 0
        self.wrong_placement = 0
        return self

    def __len__(self):
        return 1

    def __iter__(self):
        yield self

    def __repr__(self):
        text = self.last_instruction or ''
        return 'Task(instruction=%s)' % (text if len(text) < 20 else text[:20] + '...')

    def _gpu(self, grid, argmax):
        from .vec_env import max_intersection
        g = np.asarray(grid).astype(np.int8)[None]
        t = np.asarray(self.target_grid).astype(np.int8)[None]
        f = None if self.full_grid is None else np.asarray(self.full_grid).astype(np.int8)[None]
        return max_intersection(g, t, f, invariant=self.invariant, return_argmax=argmax)

    def maximal_intersection(self, grid):
        """task.py:147-161, evaluated on the GPU."""
        return int(self._gpu(grid, False)[0].item())

    def argmax_intersection(self, grid):
        """task.py:121-136 -> (dx, dz, rot)"""
        _, arg = self._gpu(grid, True)
        return tuple(int(v) for v in arg[0].tolist())

    def get_intersection(self, grid, dx, dz, rot):
        """task.py:138-145: intersection of `grid` with rotation `rot` of the target shifted by (dx, dz).
        Evaluated on the GPU as a single-shift (invariant=False) maximal_intersection of the rotated target
        moved by (-dx, -dz); target cells that the shift pushes outside the zone are clipped as in the
        reference's slicing."""
        from .vec_env import max_intersection
        tk = np.rot90(np.asarray(self.target_grid), k=-rot, axes=(1, 2))
        moved = np.zeros_like(tk)
        xs, xd = (slice(max(dx, 0), 11 + min(dx, 0)), slice(max(-dx, 0), 11 + min(-dx, 0)))
        zs, zd = (slice(max(dz, 0), 11 + min(dz, 0)), slice(max(-dz, 0), 11 + min(-dz, 0)))
        moved[:, xd, zd] = tk[:, xs, zs]
        g = np.asarray(grid).astype(np.int8)[None]
        return int(max_intersection(g, moved.astype(np.int8)[None], invariant=False)[0].item())

    def step_intersection(self, grid):
        """task.py:103-119: (right_placement, wrong_placement, done) of one step; the intersection is only
        re-evaluated (on the GPU) when the number of blocks changed -- the reference's gate."""
        blocks = int(np.count_nonzero(np.asarray(grid)))
        self.wrong_placement = self.prev_grid_size - blocks
        self.prev_grid_size = blocks
        fresh = self.maximal_intersection(grid) if self.wrong_placement else self.max_int
        self.right_placement, self.max_int = fresh - self.max_int, fresh
        return self.right_placement, self.wrong_placement, fresh == self.target_size


class Tasks:
    """Base of task collections + the sparse <-> dense converters (interface of reference
    gridworld/tasks/task.py:163-205). Written from the layout contract in include/iglu_b200.h: the world block
    (x, y, z) lives at dense index [y + 1, x + 5, z + 5] of a 9 x 11 x 11 grid."""

    _OFFSET = np.array([BUILD_ZONE_SIZE_X // 2, 1, BUILD_ZONE_SIZE_Z // 2])  # added to (x, y, z)

    @classmethod
    def to_dense(cls, blocks):
        """Block list [(x, y, z, colour), ...] -> dense grid; anything that is not a list / tuple of block
        rows (e.g. an ndarray that is dense already) is handed back unchanged."""
        if not isinstance(blocks, (list, tuple)) or not all(isinstance(b, (list, tuple)) for b in blocks):
            return blocks
        grid = np.zeros(BUILD_ZONE_SIZE, dtype=int)
        if len(blocks):
            rows = np.asarray(blocks, dtype=np.int64).reshape(-1, 4)
            x, y, z = (rows[:, :3] + cls._OFFSET).T
            flat = np.ravel_multi_index((y, x, z), BUILD_ZONE_SIZE)
            # a cell listed twice keeps its LAST colour: np.unique on the reversed order finds the last mention
            _, rev = np.unique(flat[::-1], return_index=True)
            keep = len(flat) - 1 - rev
            grid.flat[flat[keep]] = rows[keep, 3]
        return grid

    @classmethod
    def to_sparse(cls, blocks):
        """ndarray -> block list. NB the reference reads axis 0 of the array as x (task.py:177-187) although
        to_dense writes y there; a drop-in has to give the same rows, so (i0 - 5, i1 - 1, i2 - 5, value)."""
        if not isinstance(blocks, np.ndarray):
            return blocks
        idx = np.argwhere(blocks)
        vals = blocks[tuple(idx.T)]
        shifted = idx - cls._OFFSET
        return [(*(int(v) for v in r), int(c)) for r, c in zip(shifted, vals)]


# ---------------------------------------------------------------------------------------------------
# Feeding task collections to the device-resident task generator (IGLUVecEnv.set_task_pool)

def _is_task(obj):
    """A leaf task (this package's Task, the reference's Task, or anything shaped like one): it carries its
    own target_grid. Collections forward attribute look-ups to their current task, hence vars() not hasattr()."""
    return isinstance(obj, Task) or 'target_grid' in getattr(obj, '__dict__', {})


def enumerate_tasks(tasks, samples=0):
    """All task objects of a task / task collection, in a stable order. Collections are recognised by shape, not
    by class, so the reference's own objects can be passed: a `.tasks` dict or list (CustomTasks, RandomTasks
    with max_cache > 0 -- whose __iter__ yields dict KEYS in the reference, task_set.py:107-109), an iterable of
    tasks (Subtasks yields one task per turn, task.py:251-253; plain lists), a single task. A generator without
    a cache (RandomTasks(max_cache=0)) is sampled `samples` times through its sample_task()."""
    if _is_task(tasks):
        return [tasks]
    inner = getattr(tasks, '__dict__', {}).get('tasks')
    if isinstance(inner, dict):
        inner = list(inner.values())
    if isinstance(inner, (list, tuple)) and len(inner):
        return [t for v in inner for t in enumerate_tasks(v)]
    if inner is not None and callable(getattr(tasks, 'sample_task', None)):
        return [tasks.sample_task() for _ in range(int(samples))]
    return [t for v in tasks for t in enumerate_tasks(v)]


def _dense_int8(grid_or_blocks):
    """Dense int8 grid of a starting / target grid given either as a block list or as a dense array."""
    if grid_or_blocks is None:
        return np.zeros(BUILD_ZONE_SIZE, np.int8)
    if isinstance(grid_or_blocks, np.ndarray) and grid_or_blocks.shape == BUILD_ZONE_SIZE:
        return grid_or_blocks.astype(np.int8)
    rows = [tuple(int(v) for v in b) for b in grid_or_blocks]
    return np.asarray(Tasks.to_dense(rows)).astype(np.int8)


def pool_arrays(tasks, samples=0):
    """(targets int8 [m, 9, 11, 11], starts int8 [m, 9, 11, 11], chats list[str]) of a task collection: the
    inputs of IGLUVecEnv.set_task_pool (starting_grid None is treated as [] like everywhere in this package;
    starting grids may be block lists or dense arrays)."""
    ts = enumerate_tasks(tasks, samples)
    if not ts:
        raise ValueError('empty task collection')
    targets = np.stack([_dense_int8(np.asarray(t.target_grid)) for t in ts])
    starts = np.stack([_dense_int8(t.starting_grid) for t in ts])
    chats = [getattr(t, 'chat', '') for t in ts]
    return targets, starts, chats


def staged_arrays(structures):
    """Arrays for IGLUVecEnv.set_staged_tasks from multi-turn structures -- objects with a `.structure_seq` (the shape
    of the reference's Subtasks, tasks/task.py:208-308; `.dialog` is used for the chats when present) or plain lists
    [blocks after turn 0, blocks after turn 1, ...] of (x, y, z, colour) rows.

    Task pool (what an episode can be): for a structure with L turns the rows turn = -1 .. L - 2, i.e. "start from the
    structure after turn `turn` (nothing for -1), build the one after turn + 1" -- the tasks Subtasks.reset() creates.
    Stage rows (what the goal can become): turns 0 .. L - 1 plus one repeat of the last turn (see
    include/iglu_b200.h iglu_staged_tasks), each scored against the FULL structure's admissible shifts."""
    pool_t, pool_s, chats, rows = [], [], [], []
    stage_t, stage_f, next_stage, stage_of, base = [], [], [], [], []
    for si, st in enumerate(structures):
        seq = getattr(st, 'structure_seq', st)
        dialog = getattr(st, 'dialog', None) or [''] * len(seq)
        dense = [_dense_int8([tuple(int(v) for v in b) for b in blocks]) for blocks in seq]
        L = len(dense)
        if L == 0:
            raise ValueError('a structure needs at least one turn')
        b0 = len(stage_t)
        base.append(b0)
        for j in range(L + 1):
            stage_t.append(dense[min(j, L - 1)])
            stage_f.append(dense[L - 1])
            next_stage.append(b0 + j + 1 if j < L else -1)
        lines = []
        for turn in range(-1, L - 1):
            goal = turn + 1
            if goal < len(dialog):
                utt = dialog[goal]
                lines.append('\n'.join(utt) if isinstance(utt, list) else utt)
            pool_t.append(dense[goal])
            pool_s.append(dense[turn] if turn >= 0 else np.zeros(BUILD_ZONE_SIZE, np.int8))
            chats.append('\n'.join(lines))
            rows.append((si, turn))
            stage_of.append(b0 + goal)
    return dict(pool_targets=np.stack(pool_t), pool_starts=np.stack(pool_s), chats=chats, rows=rows,
                stage_targets=np.stack(stage_t), stage_fulls=np.stack(stage_f),
                next_stage=np.array(next_stage, np.int32), stage_of=np.array(stage_of, np.int32),
                stage_base=np.array(base, np.int32))


class _Record:
    """Attribute bag standing in for a pickled reference object (its class is never imported)."""

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {})


class _StagedRecord(_Record):
    """A pickled reference Subtasks: dialog, structure_seq (block lists after every turn), ..."""


class _TaskRecord(_Record):
    """A pickled reference Task: target_grid, starting_grid, chat, ..."""


class _DatasetUnpickler(pickle.Unpickler):
    """IGLUDataset.dump_tasks_dataset (reference data/iglu_dataset.py:204-216) pickles {task id: [Subtasks, ...]}.
    The two task classes are mapped to attribute bags and apart from them only numpy array / scalar reconstruction
    and builtin containers are allowed, so that loading a file cannot run arbitrary code."""
    _ALLOWED = {('numpy.core.multiarray', 'scalar'), ('numpy._core.multiarray', 'scalar'), ('numpy', 'dtype'),
                ('numpy.core.multiarray', '_reconstruct'), ('numpy._core.multiarray', '_reconstruct'),
                ('numpy', 'ndarray'), ('builtins', 'tuple'), ('builtins', 'list'), ('builtins', 'dict'),
                ('builtins', 'set'), ('collections', 'OrderedDict'), ('collections', 'defaultdict'),
                ('copyreg', '_reconstructor'), ('builtins', 'object')}
    _CLASSES = {('gridworld.tasks.task', 'Subtasks'): _StagedRecord, ('gridworld.tasks.task', 'Task'): _TaskRecord,
                ('gridworld.tasks', 'Subtasks'): _StagedRecord, ('gridworld.tasks', 'Task'): _TaskRecord}

    def find_class(self, module, name):
        if (module, name) in self._CLASSES:
            return self._CLASSES[(module, name)]
        if (module, name) in self._ALLOWED:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f'task dataset refers to {module}.{name}: not allowed')


def _staged_tasks(rec):
    """The single-turn tasks of one staged (multi-turn) structure, as the reference enumerates them (Subtasks.__iter__
    / create_task, task.py:251-285): turn i starts from the blocks after turn i - 1 (nothing for i = 0), its target is
    the blocks after turn i, and its chat is every utterance up to and including turn i, one per line."""
    seq, dialog = rec.structure_seq, rec.dialog
    lines, out = [], []
    for i, blocks in enumerate(seq):
        if i < len(dialog):
            turn = dialog[i]
            lines.append('\n'.join(turn) if isinstance(turn, list) else turn)
        out.append(Task('\n'.join(lines), _dense_int8(list(blocks)).astype(np.int32),
                        starting_grid=[tuple(int(v) for v in b) for b in (seq[i - 1] if i else [])],
                        full_grid=_dense_int8(list(seq[-1])).astype(np.int32)))
    return out


def load_tasks_dataset(path):
    """Reads a parsed-task cache in the format of the reference's IGLUDataset.dump_tasks_dataset (bz2-compressed
    pickle of {task id: [Subtasks, ...]}, data/iglu_dataset.py:204-216) without importing the reference, and returns
    [(task id, session index, [Task per turn])]. Feed the tasks to IGLUVecEnv.set_task_generator / pool_arrays."""
    import bz2
    with open(path, 'rb') as f:
        raw = bz2.decompress(f.read())
    data = _DatasetUnpickler(io.BytesIO(raw)).load()
    if not isinstance(data, dict):
        raise ValueError('task dataset must hold a dict task id -> list of staged tasks')
    out = []
    for tid in sorted(data, key=str):
        sessions = data[tid] if isinstance(data[tid], (list, tuple)) else [data[tid]]
        for k, rec in enumerate(sessions):
            if isinstance(rec, _StagedRecord):
                out.append((str(tid), k, _staged_tasks(rec)))
            elif isinstance(rec, _TaskRecord):
                out.append((str(tid), k, [Task(rec.chat, np.asarray(rec.target_grid),
                                               starting_grid=getattr(rec, 'starting_grid', None),
                                               full_grid=getattr(rec, 'full_grid', None))]))
            else:
                raise ValueError(f'task dataset entry {tid!r} holds {type(rec).__name__}, not a task')
    return out


class _GoalsUnpickler(pickle.Unpickler):
    """goals.pkl holds {name: [(x, y, z, colour), ...]} with numpy integer scalars: only
    what is needed to rebuild those is allowed, so that loading a file cannot run arbitrary code."""
    _ALLOWED = {('numpy.core.multiarray', 'scalar'), ('numpy._core.multiarray', 'scalar'), ('numpy', 'dtype'),
                ('numpy.core.multiarray', '_reconstruct'), ('numpy._core.multiarray', '_reconstruct'),
                ('numpy', 'ndarray'), ('builtins', 'tuple'), ('builtins', 'list'), ('builtins', 'dict'),
                ('collections', 'OrderedDict')}

    def find_class(self, module, name):
        if (module, name) in self._ALLOWED:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f'goals file refers to {module}.{name}: not allowed')


def load_goals(path):
    """Reads a structure file in the format of the reference's skills/goals.pkl and returns CustomTasks-style
    goals [(name, [(x, y, z, colour), ...]), ...] sorted by name (C1, C2, ...)."""
    with open(path, 'rb') as f:
        raw = f.read()
    goals = _GoalsUnpickler(io.BytesIO(raw)).load()
    if not isinstance(goals, dict):
        raise ValueError('goals file must hold a dict name -> block list')

    def key(nm):
        digits = ''.join(ch for ch in str(nm) if ch.isdigit())
        return (int(digits) if digits else 0, str(nm))
    out = []
    for nm in sorted(goals.keys(), key=key):
        out.append((str(nm), [tuple(int(v) for v in b) for b in goals[nm]]))
    return out
