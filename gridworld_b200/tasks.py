"""Host-side task objects with the reference's interface (reference gridworld/tasks/task.py,
task_set.py). They only carry data; maximal_intersection / argmax_intersection run on the GPU
through iglu_max_intersection, and inside an env the synthetic task is prepared on the device by
iglu_set_tasks (rotations, admissible rectangles, block list)."""
import io
import pickle
import uuid

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
        ins = self.last_instruction or ''
        ins = ins if len(ins) < 20 else ins[:20] + '...'
        return f'Task(instruction={ins})'

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
        """task.py:103-119"""
        grid = np.asarray(grid)
        grid_size = int((grid != 0).sum())
        wrong_placement = self.prev_grid_size - grid_size
        max_int = self.maximal_intersection(grid) if wrong_placement != 0 else self.max_int
        done = max_int == self.target_size
        self.prev_grid_size = grid_size
        right_placement = max_int - self.max_int
        self.max_int = max_int
        self.right_placement = right_placement
        self.wrong_placement = wrong_placement
        return right_placement, wrong_placement, done


class Tasks:
    """reference gridworld/tasks/task.py:163-205"""

    @classmethod
    def to_dense(cls, blocks):
        if isinstance(blocks, (list, tuple)):
            if all(isinstance(b, (list, tuple)) for b in blocks):
                grid = np.zeros(BUILD_ZONE_SIZE, dtype=int)
                for x, y, z, block_id in blocks:
                    grid[y + 1, x + BUILD_ZONE_SIZE_X // 2, z + BUILD_ZONE_SIZE_Z // 2] = block_id
                blocks = grid
        return blocks

    @classmethod
    def to_sparse(cls, blocks):
        # NB the reference treats axis 0 of an ndarray as x here (task.py:177-187); kept as is.
        if isinstance(blocks, np.ndarray):
            idx = blocks.nonzero()
            types = [blocks[i] for i in zip(*idx)]
            blocks = [(*i, t) for *i, t in zip(*idx, types)]
            new_blocks = []
            for x, y, z, bid in blocks:
                new_blocks.append((x - BUILD_ZONE_SIZE_X // 2, y - 1, z - BUILD_ZONE_SIZE_Z // 2, bid))
            blocks = new_blocks
        return blocks

    def reset(self):
        return NotImplemented

    def __len__(self):
        return NotImplemented

    def __iter__(self):
        return NotImplemented

    def set_task(self, task_id):
        return NotImplemented

    def get_target(self):
        return NotImplemented

    def set_task_obj(self, task):
        return NotImplemented


class CustomTasks(Tasks):
    """reference gridworld/tasks/task_set.py:23-56"""

    def __init__(self, goals, task_kwargs=None):
        super().__init__()
        self.task_kwargs = task_kwargs or {}
        self.tasks = {str(uuid.uuid4().hex): Task(conv, self.to_dense(grid), **self.task_kwargs)
                      for conv, grid in goals}
        self.task_ids = list(self.tasks.keys())
        self.reset()

    def __getattr__(self, name):
        if name in ('current', 'tasks', 'task_ids', 'task_kwargs'):
            raise AttributeError(name)
        return getattr(self.current, name)

    def __len__(self):
        return len(self.task_ids)

    def __iter__(self):
        for task in self.tasks.values():
            yield from iter(task)

    def reset(self):
        task_id = np.random.choice(self.task_ids)
        self.current = self.tasks[task_id].reset()
        return self.current


class RandomTasks(Tasks):
    """reference gridworld/tasks/task_set.py:59-157 (same numpy global-RNG call order)"""

    def __init__(self, max_blocks=4, height_levels=1, allow_float=False, max_dist=2, num_colors=1, max_cache=0):
        self.height_levels = height_levels
        self.max_blocks = max_blocks
        self.allow_float = allow_float
        self.max_dist = max_dist
        self.num_colors = num_colors
        self.max_cache = max_cache
        self.tasks = {}
        self.current = None
        for _ in range(self.max_cache):
            self.tasks[str(uuid.uuid4().hex)] = self.sample_task()
        self.reset()

    def __getattr__(self, name):
        if name in ('current', 'tasks'):
            raise AttributeError(name)
        return getattr(self.current, name)

    def __len__(self):
        return self.max_cache

    def __iter__(self):
        for task in self.tasks:
            yield task

    def reset(self):
        if self.max_cache > 0:
            sample = np.random.choice(list(self.tasks.keys()))
            self.current = self.tasks[sample]
            self.current_id = sample
            return self.current
        self.current = self.sample_task()
        return self.current

    def set_task(self, task_id):
        self.current = self.tasks[task_id]
        return self.current

    def sample_task(self):
        target_grid = np.zeros(BUILD_ZONE_SIZE, dtype=np.int32)
        for height in range(self.height_levels):
            block_x = np.random.choice(BUILD_ZONE_SIZE_X)
            block_z = np.random.choice(BUILD_ZONE_SIZE_Z)
            color = np.random.choice(self.num_colors) + 1
            target_grid[height, block_x, block_z] = color
            for _ in range(self.max_blocks - 1):
                ddx, ddz = 0, 0
                while ddx == 0 and ddz == 0 \
                        or block_x + ddx >= BUILD_ZONE_SIZE_X or block_z + ddz >= BUILD_ZONE_SIZE_Z \
                        or block_x + ddx < 0 or block_z + ddz < 0 \
                        or target_grid[height, block_x + ddx, block_z + ddz] != 0:
                    ddx = np.random.choice(2 * self.max_dist + 1) - self.max_dist
                    ddz = np.random.choice(2 * self.max_dist + 1) - self.max_dist
                    color = np.random.choice(self.num_colors) + 1
                target_grid[height, block_x + ddx, block_z + ddz] = color
        return Task('', target_grid)


class Subtasks(Tasks):
    """Staged task where subtasks are separate segments (reference gridworld/tasks/task.py:208-308);
    host-side task producer with the reference's interface and numpy global-RNG call order."""

    def __init__(self, dialog, structure_seq, invariant=False, progressive=True):
        self.dialog = dialog
        self.invariant = invariant
        self.progressive = progressive
        self.structure_seq = structure_seq
        self.next = None
        self.full = False
        self.task_start = 0
        self.task_goal = 0
        self.full_structure = self.to_dense(self.structure_seq[-1])
        self.current = self.reset()

    def __getattr__(self, name):
        if name == 'current':
            return None
        return getattr(self.current, name)

    def reset(self):
        if self.next is None:
            if len(self.structure_seq) == 1:
                turn = -1
            else:
                turn = np.random.choice(len(self.structure_seq)) - 1
            turn_goal = turn + 1
        else:
            turn = self.next
            turn_goal = self.next + 1
        self.task_start = turn
        self.task_goal = turn_goal
        self.current = self.create_task(self.task_start, self.task_goal)
        return self.current

    def __len__(self):
        return len(self.structure_seq)

    def __iter__(self):
        for i in range(len(self)):
            yield self.create_task(i - 1, i)

    def __repr__(self):
        return (f"Subtasks(total_steps={len(self.structure_seq)}, "
                f"current_task_start={self.task_start}, current_task_end={self.task_goal})")

    def create_task(self, turn_start, turn_goal):
        dialog = ''
        for turn in self.dialog[:turn_goal + 1]:
            if isinstance(turn, list):
                turn = '\n'.join(turn)
            dialog += '\n' + turn if len(dialog) > 0 else turn
        initial_blocks = [] if turn_start == -1 else self.structure_seq[turn_start]
        tid = min(turn_goal, len(self.structure_seq) - 1) if not self.full else -1
        target_grid = self.structure_seq[tid]
        task = Task(dialog, target_grid=self.to_dense(target_grid), starting_grid=self.to_sparse(initial_blocks),
                    full_grid=self.full_structure, last_instruction='\n'.join(self.dialog[tid]))
        task.reset()
        return task

    def step_intersection(self, grid):
        right_placement, wrong_placement, done = self.current.step_intersection(grid)
        if done and len(self.structure_seq) > self.task_goal and self.progressive:
            self.task_goal += 1
            self.current = self.create_task(self.task_start, self.task_goal)
            self.current.prev_grid_size = 0
            _, _, done = self.current.step_intersection(grid)
        return right_placement, wrong_placement, done

    def set_task(self, task_id):
        self.task_id = task_id
        self.current = self.create_task(task_id - 1, task_id)
        return self.current

    def set_task_obj(self, task):
        self.task_id = None
        self.current = task
        return self.current


# ---------------------------------------------------------------------------------------------------
# Feeding task collections to the device-resident task generator (IGLUVecEnv.set_task_pool)

def enumerate_tasks(tasks):
    """All Task objects of a Task / Tasks-like collection, in a stable order.

    Handles the reference's container shapes: objects with a `.tasks` dict or list (CustomTasks, RandomTasks with
    max_cache > 0 -- whose __iter__ yields dict KEYS in the reference, task_set.py:107-109), Subtasks (one Task
    per turn, task.py:251-253), a plain Task, or any iterable of Tasks."""
    if isinstance(tasks, Task):
        return [tasks]
    inner = getattr(tasks, '__dict__', {}).get('tasks')
    if isinstance(inner, dict):
        return [t for v in inner.values() for t in enumerate_tasks(v)]
    if isinstance(inner, (list, tuple)):
        return [t for v in inner for t in enumerate_tasks(v)]
    out = []
    for t in tasks:
        out.extend(enumerate_tasks(t) if not isinstance(t, Task) else [t])
    return out


def pool_arrays(tasks):
    """(targets int8 [m, 9, 11, 11], starts int8 [m, 9, 11, 11], chats list[str]) of a task collection: the
    inputs of IGLUVecEnv.set_task_pool (starting_grid None is treated as [] like everywhere in this package)."""
    ts = enumerate_tasks(tasks)
    if not ts:
        raise ValueError('empty task collection')
    targets = np.zeros((len(ts), *BUILD_ZONE_SIZE), np.int8)
    starts = np.zeros((len(ts), *BUILD_ZONE_SIZE), np.int8)
    chats = []
    for i, t in enumerate(ts):
        targets[i] = np.asarray(t.target_grid).astype(np.int8)
        sg = t.starting_grid
        if sg is not None and len(sg) > 0:
            starts[i] = np.asarray(Tasks.to_dense([tuple(int(v) for v in b) for b in sg])).astype(np.int8)
        chats.append(t.chat)
    return targets, starts, chats


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
