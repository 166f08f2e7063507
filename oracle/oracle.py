"""ctypes binding of oracle/libiglu_oracle.so (C restatement of the reference step path).

TEST INFRASTRUCTURE ONLY -- see oracle/iglu_oracle.c for the reference file:line citations.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, 'libiglu_oracle.so')
CELLS = 9 * 11 * 11


def build(force=False):
    src = os.path.join(_HERE, 'iglu_oracle.c')
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, '-B', 'libiglu_oracle.so'], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, i32, i64, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_double
        L.orc_create.restype = vp
        L.orc_create.argtypes = [i32, i32, i32, i32, f64, f64]
        L.orc_destroy.argtypes = [vp]
        L.orc_set_task.argtypes = [vp, i32, vp, vp, vp]
        L.orc_reset.argtypes = [vp, vp]
        L.orc_set_pool.argtypes = [vp, vp, vp, i64, C.c_uint64, i64]
        L.orc_pool_next_task.restype = C.POINTER(C.c_int32)
        L.orc_pool_next_task.argtypes = [vp]
        L.orc_pool_task_id.restype = C.POINTER(C.c_int32)
        L.orc_pool_task_id.argtypes = [vp]
        L.orc_get_target.argtypes = [vp, vp]
        L.orc_step_walking.argtypes = [vp, vp, vp, vp, i32, i32]
        L.orc_step_flying.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32]
        L.orc_step_walking_dict.argtypes = [vp, vp, vp, vp, vp, i32, i32]
        L.orc_get_state.argtypes = [vp] + [vp] * 10
        L.orc_set_agent.argtypes = [vp, i32, vp, vp, f64, i32, i32]
        L.orc_get_obs.argtypes = [vp, vp, vp, vp, vp]
        L.orc_max_intersection.argtypes = [vp, vp, vp, i32, vp, vp, i64, i32]
        L.orc_task_describe.argtypes = [vp, vp, i32, vp, vp, vp]
        L.orc_tasks_prepare.restype = vp
        L.orc_tasks_prepare.argtypes = [vp, i64, i32]
        L.orc_tasks_eval.argtypes = [vp, vp, vp, i64, i32]
        L.orc_tasks_free.argtypes = [vp]
        L.orc_staged_create.restype = vp
        L.orc_staged_create.argtypes = [vp, i32, i32]
        L.orc_staged_destroy.argtypes = [vp]
        L.orc_staged_reset.argtypes = [vp, i32]
        L.orc_staged_step.argtypes = [vp, vp, vp]
        L.orc_sample_tasks.argtypes = [vp, i64, i32, i32, i32, i32, C.c_uint64, i64]
        L.orc_render_pov.argtypes = [vp, vp, i32, i32, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i8(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(np.asarray(a), dtype=np.int8)
    return a if shape is None else a.reshape(shape)


class OracleBatch:
    """N independent reference envs (render=False, vector_state=True) stepped in lock-step."""

    def __init__(self, n, action_space='walking', select_and_place=True, max_steps=250,
                 right_placement_scale=1.0, wrong_placement_scale=0.1, threads=1):
        self.n = int(n)
        self.flying = action_space == 'flying'
        self.threads = int(threads)
        self._h = lib().orc_create(self.n, int(self.flying), int(bool(select_and_place)), int(max_steps),
                                   float(right_placement_scale), float(wrong_placement_scale))

    def __del__(self):
        if getattr(self, '_h', None) and _lib is not None:
            _lib.orc_destroy(self._h)
            self._h = None

    def set_task(self, i, target, start=None, pose=None):
        t = _i8(target, (CELLS,))
        s = _i8(start, (CELLS,))
        p = None if pose is None else np.ascontiguousarray(pose, dtype=np.float64)
        lib().orc_set_task(self._h, int(i), _p(t), _p(s), _p(p))

    def set_tasks(self, targets, starts=None, poses=None):
        targets = _i8(targets).reshape(self.n, CELLS)
        starts = None if starts is None else _i8(starts).reshape(self.n, CELLS)
        for i in range(self.n):
            self.set_task(i, targets[i], None if starts is None else starts[i],
                          None if poses is None else poses[i])

    def set_pool(self, targets, starts=None, seed=0, env_offset=0):
        """env.set_task_generator with a finite pool drawn at every reset (see iglu_oracle.c batch_reset_env)."""
        t = _i8(targets)
        m = t.shape[0]
        t = t.reshape(m, CELLS)
        s = None if starts is None else _i8(starts).reshape(m, CELLS)
        lib().orc_set_pool(self._h, _p(t), _p(s), m, int(seed) & (2 ** 64 - 1), int(env_offset))
        self.next_task = np.ctypeslib.as_array(lib().orc_pool_next_task(self._h), shape=(self.n,))
        self.task_ids = np.ctypeslib.as_array(lib().orc_pool_task_id(self._h), shape=(self.n,))

    def target_grids(self):
        out = np.zeros((self.n, 9, 11, 11), np.int32)
        lib().orc_get_target(self._h, _p(out))
        return out

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        lib().orc_reset(self._h, _p(m))

    def set_agent(self, i, pos=None, rot=None, dy=0.0, tis=2, active=1):
        p = None if pos is None else np.ascontiguousarray(pos, dtype=np.float64)
        r = None if rot is None else np.ascontiguousarray(rot, dtype=np.float64)
        lib().orc_set_agent(self._h, int(i), _p(p), _p(r), float(dy), int(tis), int(active))

    def _out(self):
        return np.zeros(self.n, np.float64), np.zeros(self.n, np.uint8)

    def step_walking(self, actions, auto_reset=False):
        a = np.ascontiguousarray(actions, dtype=np.int32).reshape(self.n)
        r, d = self._out()
        lib().orc_step_walking(self._h, _p(a), _p(r), _p(d), int(auto_reset), self.threads)
        return r, d.astype(bool)

    def step_flying(self, movement, camera, inventory, placement, auto_reset=False):
        mv = np.ascontiguousarray(movement, dtype=np.float32).reshape(self.n, 3)
        cam = np.ascontiguousarray(camera, dtype=np.float32).reshape(self.n, 2)
        inv = np.ascontiguousarray(inventory, dtype=np.int32).reshape(self.n)
        pl = np.ascontiguousarray(placement, dtype=np.int32).reshape(self.n)
        r, d = self._out()
        lib().orc_step_flying(self._h, _p(mv), _p(cam), _p(inv), _p(pl), _p(r), _p(d), int(auto_reset),
                              self.threads)
        return r, d.astype(bool)

    def step_walking_dict(self, buttons, camera, auto_reset=False):
        b = np.ascontiguousarray(buttons, dtype=np.int32).reshape(self.n, 8)
        cam = np.ascontiguousarray(camera, dtype=np.float32).reshape(self.n, 2)
        r, d = self._out()
        lib().orc_step_walking_dict(self._h, _p(b), _p(cam), _p(r), _p(d), int(auto_reset), self.threads)
        return r, d.astype(bool)

    def state(self):
        n = self.n
        s = dict(grid=np.zeros((n, 9, 11, 11), np.int8), pos=np.zeros((n, 3)), rot=np.zeros((n, 2)),
                 dy=np.zeros(n), tis=np.zeros(n, np.int32), inventory=np.zeros((n, 6), np.int32),
                 active=np.zeros(n, np.int32), step_no=np.zeros(n, np.int32), max_int=np.zeros(n, np.int32),
                 err=np.zeros(n, np.int32))
        lib().orc_get_state(self._h, *[_p(s[k]) for k in ('grid', 'pos', 'rot', 'dy', 'tis', 'inventory',
                                                          'active', 'step_no', 'max_int', 'err')])
        return s

    def obs(self):
        n = self.n
        o = dict(inventory=np.zeros((n, 6), np.float32), compass=np.zeros((n, 1), np.float32),
                 agentPos=np.zeros((n, 5), np.float32), grid=np.zeros((n, 9, 11, 11), np.int32))
        lib().orc_get_obs(self._h, _p(o['inventory']), _p(o['compass']), _p(o['agentPos']), _p(o['grid']))
        return o


def max_intersection(grids, targets, fulls=None, invariant=True, threads=1, return_argmax=False):
    """Task(target, full_grid=full, invariant=invariant).maximal_intersection(grid) per pair."""
    g = _i8(grids).reshape(-1, CELLS)
    t = _i8(targets).reshape(-1, CELLS)
    f = None if fulls is None else _i8(fulls).reshape(-1, CELLS)
    n = g.shape[0]
    out = np.zeros(n, np.int32)
    arg = np.zeros((n, 3), np.int32) if return_argmax else None
    lib().orc_max_intersection(_p(g), _p(t), _p(f), int(bool(invariant)), _p(out), _p(arg), n, int(threads))
    return (out, arg) if return_argmax else out


def task_describe(target, full=None, invariant=True):
    t = _i8(target, (CELLS,))
    f = _i8(full, (CELLS,))
    n_adm = np.zeros(4, np.int32)
    rect = np.zeros((4, 4), np.int32)
    ts = np.zeros(1, np.int32)
    lib().orc_task_describe(_p(t), _p(f), int(bool(invariant)), _p(n_adm), _p(rect), _p(ts))
    return n_adm, rect, int(ts[0])


class PreparedTasks:
    """Tasks built once (Task.__init__), then only maximal_intersection is evaluated (timing helper)."""

    def __init__(self, targets, threads=1):
        t = _i8(targets).reshape(-1, CELLS)
        self.n = t.shape[0]
        self.threads = int(threads)
        self._h = lib().orc_tasks_prepare(_p(t), self.n, self.threads)

    def eval(self, grids):
        g = _i8(grids).reshape(self.n, CELLS)
        out = np.zeros(self.n, np.int32)
        lib().orc_tasks_eval(self._h, _p(g), _p(out), self.n, self.threads)
        return out

    def __del__(self):
        if getattr(self, '_h', None) and _lib is not None:
            _lib.orc_tasks_free(self._h)
            self._h = None


def sample_tasks(m, max_blocks=4, height_levels=1, max_dist=2, num_colors=1, seed=0, first_task=0):
    """RandomTasks.sample_task (task_set.py:135-157) for m tasks with the batched build's counter-based stream."""
    out = np.zeros((int(m), 9, 11, 11), np.int8)
    lib().orc_sample_tasks(_p(out), int(m), int(max_blocks), int(height_levels), int(max_dist), int(num_colors),
                           int(seed) & (2 ** 64 - 1), int(first_task))
    return out


class Staged:
    """Subtasks (reference tasks/task.py:208-308) for one multi-turn structure: reset(turn) = Subtasks.reset with
    self.next = turn; step(grid) = Subtasks.step_intersection(grid) -> (right, wrong, done, task_goal, max_int)."""

    def __init__(self, stages, progressive=True):
        st = _i8(stages).reshape(-1, CELLS)
        self._h = lib().orc_staged_create(_p(st), st.shape[0], int(bool(progressive)))

    def reset(self, turn):
        lib().orc_staged_reset(self._h, int(turn))

    def step(self, grid):
        g = _i8(grid, (CELLS,))
        out = np.zeros(5, np.int32)
        lib().orc_staged_step(self._h, _p(g), _p(out))
        return tuple(int(v) for v in out)

    def __del__(self):
        if getattr(self, '_h', None) and _lib is not None:
            _lib.orc_staged_destroy(self._h)
            self._h = None


def render_pov(grid, pose, height=64, width=64, atlas=None):
    """obs['pov'] of one env (orc_render_pov: the ray-cast restatement of gridworld/render.py). grid: int8 [9, 11, 11];
    pose: (x, y, z, yaw, pitch); atlas: uint8 [128, 128, 4] (texture.png, rows top-down) or None = flat palette."""
    g = _i8(grid, (CELLS,))
    p = np.ascontiguousarray(np.asarray(pose, np.float64).reshape(5))
    a = None if atlas is None else np.ascontiguousarray(np.asarray(atlas, np.uint8).reshape(128, 128, 4))
    out = np.zeros((int(height), int(width), 3), np.uint8)
    lib().orc_render_pov(_p(g), _p(p), int(height), int(width), None if a is None else _p(a), _p(out))
    return out
