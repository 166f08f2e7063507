"""Batched IGLU Gridworld: N environments resident in HBM, stepped by the sm_100a kernels behind
the C ABI of include/iglu_b200.h.

This is the batched form of the reference's GridWorld (reference gridworld/env.py:26-303) with
render=False, vector_state=True. PyTorch only hands out the device tensors; every state update,
observation and reward is computed by libiglu_b200.so (there is no CPU / eager fallback).
"""
import contextlib
import ctypes as C

import numpy as np
import torch

from . import _lib
from .spaces import Box, Dict, Discrete, String

_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)

_ACTION_SPACES = {'walking': None, 'flying': _lib.FLYING}


def _require_cuda(device):
    if not torch.cuda.is_available():
        raise RuntimeError('gridworld_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    dev = torch.device(device)
    if dev.type != 'cuda':
        raise RuntimeError(f'gridworld_b200 only runs on CUDA devices, got {device!r}')
    if dev.index is None:
        dev = torch.device('cuda', torch.cuda.current_device())
    return dev


def make_action_space(action_space, discretize):
    """reference gridworld/env.py:56-78"""
    if action_space == 'walking':
        if discretize:
            return Discrete(18)
        return Dict({
            'forward': Discrete(2), 'back': Discrete(2), 'left': Discrete(2), 'right': Discrete(2),
            'jump': Discrete(2), 'attack': Discrete(2), 'use': Discrete(2),
            'camera': Box(low=-5, high=5, shape=(2,)), 'hotbar': Discrete(7)})
    if action_space == 'flying':
        return Dict({
            'movement': Box(low=-1, high=1, shape=(3,), dtype=np.float32),
            'camera': Box(low=-5, high=5, shape=(2,), dtype=np.float32),
            'inventory': Discrete(7), 'placement': Discrete(3)})
    raise ValueError(f'unknown action_space {action_space!r}')


def make_observation_space(vector_state, target_in_obs, render=False, render_size=(64, 64)):
    """reference gridworld/env.py:79-95"""
    sp = {'inventory': Box(low=0, high=20, shape=(6,), dtype=np.float32),
          'compass': Box(low=-180, high=180, shape=(1,), dtype=np.float32),
          'dialog': String()}
    if vector_state:
        sp['agentPos'] = Box(low=np.array([-8, -2, -8, -90, 0], dtype=np.float32),
                             high=np.array([8, 12, 8, 90, 360], dtype=np.float32), shape=(5,))
        sp['grid'] = Box(low=-1, high=7, shape=(9, 11, 11), dtype=np.int32)
    if target_in_obs:
        sp['target_grid'] = Box(low=-1, high=7, shape=(9, 11, 11), dtype=np.int32)
    if render:
        sp['pov'] = Box(low=0, high=255, shape=(*render_size, 3), dtype=np.uint8)
    return Dict(sp)


class IGLUVecEnv:
    """N independent IGLU Gridworld envs on one GPU.

    kwargs mirror reference create_env / GridWorld.__init__ (env.py:27-31, 333-338). Additions:
      num_envs, device, auto_reset (an env that returned done is reset at the start of the next
      step call and then stepped -- per env this is exactly `reset(); step(a)` on the reference),
      reward_mode ('on_change' = the reference's gate, 'every_step' = evaluate maximal_intersection
      on every step; outputs are identical), grid_obs (True: rewrite the int32 obs['grid'] in full every step;
      'incremental': keep the persistent obs['grid'] buffer up to date cell by cell -- same contents, ~20x
      fewer bytes; False: no int32 copy, obs['grid'] is the int8 state view),
      alternate_phases (scheduling only: in the fused launch half of the warps write observations before the
      physics), split_launch (scheduling only: True = the library's measured policy -- the overlapped step (widen next
      to the physics, then commit) for a fully rewritten obs['grid'] on batches >= 8,192 envs, otherwise physics and
      observation/reward as two launches or one fused launch; False = always one fused launch; 'force' = always two), max_target_blocks (capacity of the per-env target block list),
      render / render_size / texture (render=True adds obs['pov'], uint8 [n, h, w, 3]: the reference's OpenGL frame
      (render.py) restated as a voxel ray caster, one more launch per step / reset; texture = the reference's
      128 x 128 RGBA atlas as an array [128, 128, 4], a PNG path, or 'auto' = the texture.png of an installed reference
      package if there is one (load_texture); None = flat colours per block id).

    step() returns device tensors that alias persistent buffers: a policy reads them in place.
    set_tasks / set_task_pool / set_task_generator assign tasks (the latter two keep a device-resident generator
    that every reset draws from); step_host is the host-buffer form of step (actions in, reward / done out).
    """

    def __init__(self, num_envs, device='cuda', action_space='walking', discretize=True,
                 select_and_place=True, max_steps=250, right_placement_scale=1., wrong_placement_scale=0.1,
                 target_in_obs=False, vector_state=True, auto_reset=True, reward_mode='on_change',
                 grid_obs=True, render=False, max_target_blocks=_lib.TLIST_MAX, alternate_phases=None, split_launch=True,
                 render_size=(64, 64), texture=None, _sched_bits=0):
        if reward_mode not in ('on_change', 'every_step'):
            raise ValueError("reward_mode must be 'on_change' or 'every_step'")
        self.device = _require_cuda(device)
        self.num_envs = n = int(num_envs)
        if n <= 0:
            raise ValueError('num_envs must be positive')
        self.action_space_type = action_space
        self.discretize = bool(discretize)
        self.action_space = make_action_space(action_space, discretize)
        self.do_render = bool(render)
        self.render_size = (int(render_size[0]), int(render_size[1]))   # (width, height) as in the reference (env.py:109-111)
        self.observation_space = make_observation_space(vector_state, target_in_obs, self.do_render,
                                                        (self.render_size[1], self.render_size[0]))
        self.vector_state = bool(vector_state)
        self.target_in_obs = bool(target_in_obs)
        self.select_and_place = bool(select_and_place)
        self.max_steps = int(max_steps)
        self.right_placement_scale = right_placement_scale
        self.wrong_placement_scale = wrong_placement_scale
        self.auto_reset = bool(auto_reset)
        self.reward_mode = reward_mode
        if grid_obs not in (True, False, 'incremental'):
            raise ValueError("grid_obs must be True, False or 'incremental'")
        self.grid_obs_mode = 0 if (not grid_obs or not vector_state) else (2 if grid_obs == 'incremental' else 1)
        self.grid_obs = self.grid_obs_mode != 0
        if alternate_phases is None:  # measured: helps when few envs are evaluated per step (profiles/)
            alternate_phases = reward_mode != 'every_step'
        if action_space == 'flying':
            kind = _lib.FLYING
        else:
            kind = _lib.WALKING_DISCRETE if discretize else _lib.WALKING_DICT
        self._kind = kind
        self._L = _lib.load()
        self._tasks_set = False
        self._was_reset = False   # every env has been reset since its task rows were (re)written
        self._reset_seen = None
        self._pool = None
        self._staged = None
        self.dialogs = None

        dev = self.device
        z = lambda *shape, dtype: torch.zeros(*shape, dtype=dtype, device=dev)  # noqa: E731
        # ---- state (layout documented in include/iglu_b200.h) ----
        self._grid = z(n, _lib.GRID_STRIDE, dtype=torch.int8)
        self._init = z(n, _lib.GRID_STRIDE, dtype=torch.int8)
        self._target = z(n, _lib.GRID_STRIDE, dtype=torch.int8)
        self._pose = z(6, n, dtype=torch.float64)
        self._init_pose = z(5, n, dtype=torch.float64)
        self._istate = z(8, n, dtype=torch.int32)
        tlist_cap = _lib.TLIST_HDR + max(64, (int(max_target_blocks) + 7) // 8 * 8)
        if tlist_cap > _lib.TLIST_ROW_MAX:
            raise ValueError(f'max_target_blocks must be <= {_lib.TLIST_MAX}')
        self._tlist = z(n, tlist_cap, dtype=torch.int16)
        self._tmeta = z(8, n, dtype=torch.int32)
        self._sched = z(8, dtype=torch.int32)
        self._worklist = z(8, n, dtype=torch.int32)   # scratch of the overlapped step
        # ---- observations / results ----
        self._obs_grid = z(n, 9, 11, 11, dtype=torch.int32) if self.grid_obs else None
        self._obs_target = z(n, 9, 11, 11, dtype=torch.int32) if self.target_in_obs else None
        self._obs_pos = z(n, 5, dtype=torch.float32)
        self._obs_inv = z(n, 6, dtype=torch.float32)
        self._obs_compass = z(n, 1, dtype=torch.float32)
        self._reward = z(n, dtype=torch.float64)
        self._done = z(n, dtype=torch.uint8)
        self._actions_dev = z(n, dtype=torch.int32)
        self._pov = z(n, self.render_size[1], self.render_size[0], 3, dtype=torch.uint8) if self.do_render else None
        self._atlas = None
        if self.do_render and isinstance(texture, str):
            texture = load_texture(texture)
        if self.do_render and texture is not None:
            a = torch.as_tensor(np.ascontiguousarray(np.asarray(texture, dtype=np.uint8)))
            if tuple(a.shape) != (128, 128, 4):
                raise ValueError('texture must be the 128 x 128 RGBA atlas as a uint8 array [128, 128, 4]')
            self._atlas = a.to(dev)

        p = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        self._st = _lib.State(n, p(self._grid), p(self._init), p(self._target), p(self._pose),
                              p(self._init_pose), p(self._istate), p(self._tlist), p(self._tmeta), p(self._sched),
                              p(self._worklist))
        self._ob = _lib.Obs(p(self._obs_grid), p(self._obs_target), p(self._obs_pos), p(self._obs_inv),
                            p(self._obs_compass), p(self._reward), p(self._done))
        cfg = _lib.Config(kind, int(self.select_and_place), self.max_steps, int(self.auto_reset),
                          _lib.REWARD_EVERY_STEP if reward_mode == 'every_step' else _lib.REWARD_ON_CHANGE,
                          int(self.grid_obs_mode), int(tlist_cap), (0 if alternate_phases else 1) | (0 if split_launch else 2) | (4 if split_launch == 'force' else 0)
                          | int(_sched_bits),
                          float(right_placement_scale), float(wrong_placement_scale))
        self._ctx = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(self._L.iglu_ctx_create(C.byref(cfg), C.byref(self._ctx)))
        self.gpu_launches = 0  # kernels launched by this env so far
        self._snap = None
        self._st_ref, self._ob_ref = C.byref(self._st), C.byref(self._ob)
        self._actions_ptr = self._actions_dev.data_ptr()

    def __del__(self):
        ctx = getattr(self, '_ctx', None)
        if ctx is not None and ctx.value and getattr(self, '_L', None) is not None:
            self._L.iglu_ctx_destroy(ctx)
            self._ctx = None

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        if _raw_stream is not None:  # same value as current_stream().cuda_stream, without building a Stream object
            return C.c_void_p(_raw_stream(self.device.index))
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _guard(self):
        """Makes self.device current for the C-ABI call (cheap no-op when it already is)."""
        if torch.cuda.current_device() == self.device.index:
            return contextlib.nullcontext()
        return torch.cuda.device(self.device)

    def _dev(self, x, dtype, shape):
        if isinstance(x, torch.Tensor) and x.dtype == dtype and x.device == self.device and x.is_contiguous() \
                and tuple(x.shape) == tuple(shape):
            return x
        t = torch.as_tensor(x)
        if t.dtype != dtype or t.device != self.device:
            t = t.to(device=self.device, dtype=dtype)
        t = t.contiguous()
        if tuple(t.shape) != tuple(shape):
            t = t.reshape(shape)
        return t

    def _count(self):
        self.gpu_launches += self._L.iglu_last_launch_count(self._ctx)

    # ------------------------------------------------------------------ tasks
    def set_tasks(self, targets, starts=None, init_poses=None, env_idx=None, reset=True):
        """Batched env.set_task / initialize_world (reference env.py:155-193).

        targets : int8 [k, 9, 11, 11] dense target grids (dense index [y+1, x+5, z+5])
        starts  : int8 [k, 9, 11, 11] dense starting grids or None (= starting_grid [])
        init_poses : float64 [k, 5] (x, y, z, yaw, pitch) or None
        env_idx : int [k] destination envs (default: envs 0..k-1, k must equal num_envs then)
        """
        t = torch.as_tensor(targets)
        k = t.shape[0]
        t = self._dev(t, torch.int8, (k, _lib.CELLS))
        s = None if starts is None else self._dev(starts, torch.int8, (k, _lib.CELLS))
        ip = None if init_poses is None else self._dev(init_poses, torch.float64, (k, 5))
        idx = None
        if env_idx is None:
            if k != self.num_envs:
                raise ValueError('set_tasks without env_idx needs one target per env')
        else:
            idx = self._dev(env_idx, torch.int32, (k,))
        pp = lambda a: None if a is None else a.data_ptr()  # noqa: E731
        self._check_block_capacity(t, s)  # refuse BEFORE anything is written to the env's task rows
        with torch.cuda.device(self.device):
            _lib.check(self._L.iglu_set_tasks(self._ctx, C.byref(self._st), C.byref(self._ob), pp(t), pp(s),
                                              pp(ip), pp(idx), k, self._stream()))
            self._count()
        self._tasks_set = True
        if idx is None:
            self._was_reset = False  # every env got a new task: a reset is due before the next step
        if reset:
            mask = None
            if idx is not None:
                mask = torch.zeros(self.num_envs, dtype=torch.uint8, device=self.device)
                mask[idx.long()] = 1
            return self.reset(mask)
        return None

    def set_task_pool(self, targets, starts=None, seed=0, env_offset=0, next_task=None, reset=True):
        """Batched env.set_task_generator (reference env.py:166-172) with a device-resident generator.

        The generator is a finite pool of tasks drawn uniformly at every reset, i.e. what
        RandomTasks(max_cache=m).reset (task_set.py:120-127) or any Tasks-like object enumerated into
        arrays does -- but the draw, the Task.__init__ products and the copy into the env happen on the
        device inside reset / the auto-reset of step, so the host is not on the reset path.

        targets : int8 [m, 9, 11, 11]; starts : int8 [m, 9, 11, 11] or None
        seed, env_offset : the counter-based draw is a function of (seed, env_offset + env, episode number)
                           (include/iglu_b200.h), so shards of one job reproduce the single-device sequence
        next_task : optional int32 [num_envs] device tensor; an entry 0 <= j < m is used instead of the draw
                    at that env's next reset and then set to -1 (e.g. to replay a reference generator's choices)
        Pass targets=None to detach the generator.
        """
        if targets is None:
            with self._guard():
                _lib.check(self._L.iglu_task_pool_attach(self._ctx, None))
            self._pool = None
            return None
        t = torch.as_tensor(targets)
        m = t.shape[0]
        t = self._dev(t, torch.int8, (m, _lib.CELLS))
        s = None if starts is None else self._dev(starts, torch.int8, (m, _lib.CELLS))
        self._check_block_capacity(t, s)  # refuse before the ctx learns about any new buffer
        dev = self.device
        z = lambda *shape, dtype: torch.zeros(*shape, dtype=dtype, device=dev)  # noqa: E731
        bufs = dict(target=z(m, _lib.GRID_STRIDE, dtype=torch.int8), init=z(m, _lib.GRID_STRIDE, dtype=torch.int8),
                    tlist=z(m, self._tlist.shape[1], dtype=torch.int16), tmeta=z(8, m, dtype=torch.int32),
                    cursor=z(2, self.num_envs, dtype=torch.int32))
        bufs['cursor'][0] = -1
        if next_task is not None:
            if not (isinstance(next_task, torch.Tensor) and next_task.dtype == torch.int32 and next_task.device == dev
                    and next_task.is_contiguous() and tuple(next_task.shape) == (self.num_envs,)):
                raise ValueError('next_task must be a contiguous int32 device tensor of shape [num_envs]')
            bufs['next_task'] = next_task
        pool = _lib.TaskPool(m, bufs['target'].data_ptr(), bufs['init'].data_ptr(), bufs['tlist'].data_ptr(),
                             bufs['tmeta'].data_ptr(), bufs['cursor'].data_ptr(),
                             None if next_task is None else next_task.data_ptr(), int(seed) & (2 ** 64 - 1),
                             int(env_offset))
        with self._guard():
            _lib.check(self._L.iglu_task_pool_prepare(self._ctx, C.byref(pool), t.data_ptr(),
                                                      None if s is None else s.data_ptr(), self._stream()))
            self._count()
            _lib.check(self._L.iglu_task_pool_attach(self._ctx, C.byref(pool)))
        self._pool = bufs  # the ctx retains these device pointers: keep the tensors alive
        self._tasks_set = True
        self._was_reset = False
        return self.reset() if reset else None

    def set_random_task_pool(self, m, max_blocks=4, height_levels=1, max_dist=2, num_colors=1, task_seed=0,
                             first_task=0, **pool_kwargs):
        """A pool of m RandomTasks.sample_task structures (reference task_set.py:135-157) generated ON THE DEVICE
        (iglu_sample_tasks) and attached as the task generator: neither filling nor refreshing the pool involves the
        host -- call again with another task_seed / first_task to refresh it. pool_kwargs go to set_task_pool."""
        targets = sample_tasks(m, max_blocks, height_levels, max_dist, num_colors, seed=task_seed,
                               first_task=first_task, device=self.device)
        return self.set_task_pool(targets, None, **pool_kwargs)

    def set_task_generator(self, tasks, seed=0, env_offset=0, next_task=None, reset=True, samples=4096):
        """env.set_task_generator (reference env.py:166-172) for a task COLLECTION (Task, CustomTasks, RandomTasks
        with max_cache > 0, Subtasks, or any iterable of Tasks): its tasks are enumerated once into the
        device-resident pool (set_task_pool) and every reset draws one of them on the device.
        obs['dialog'] of env e is `self.dialogs[self.task_ids[e]]` (kept on the host, like the reference)."""
        from .tasks import pool_arrays
        if tasks is None:
            self.dialogs = None
            return self.set_task_pool(None)
        hp = [getattr(tasks, k, None) for k in ('max_blocks', 'height_levels', 'max_dist', 'num_colors')]
        if (all(isinstance(v, int) for v in hp) and not getattr(tasks, '__dict__', {}).get('tasks')
                and not getattr(tasks, 'max_cache', 0)):
            # a RandomTasks(max_cache=0)-shaped generator (fresh sample_task at every reset, task_set.py:120-127):
            # its structures are sampled on the device from its hyper-parameters, `samples` of them per pool
            self.dialogs = None
            return self.set_random_task_pool(samples, *hp, task_seed=seed, seed=seed, env_offset=env_offset,
                                             next_task=next_task, reset=reset)
        targets, starts, chats = pool_arrays(tasks)
        self.dialogs = chats
        return self.set_task_pool(targets, starts, seed=seed, env_offset=env_offset, next_task=next_task, reset=reset)

    def set_staged_tasks(self, structures, progressive=True, seed=0, env_offset=0, next_task=None, reset=True):
        """Multi-turn structures with progressive goals on the device (reference Subtasks, tasks/task.py:208-308).

        structures : list of objects with `.structure_seq` (+ `.dialog`), e.g. the reference's own Subtasks, or plain
                     lists of per-turn block lists (see tasks.staged_arrays).
        Every reset draws one (structure, starting turn) task from the device-resident pool (as set_task_pool does;
        next_task imposes the row -- rows are listed in self.staged_rows as (structure index, turn)). After every
        step() the staged score of Subtasks.step_intersection(env.grid) -- right / wrong placement and done of the
        CURRENT turn's task, which moves on to the next turn when complete -- is in self.staged (device tensors)."""
        from .tasks import staged_arrays
        a = staged_arrays(structures)
        self.dialogs = a['chats']
        self.staged_rows = a['rows']
        self.set_task_pool(a['pool_targets'], a['pool_starts'], seed=seed, env_offset=env_offset, next_task=next_task,
                           reset=False)
        dev, n, ms = self.device, self.num_envs, a['stage_targets'].shape[0]
        z = lambda *shape: torch.zeros(*shape, dtype=torch.int32, device=dev)  # noqa: E731
        bufs = dict(tlist=torch.zeros(ms, self._tlist.shape[1], dtype=torch.int16, device=dev), tmeta=z(8, ms),
                    next_stage=torch.as_tensor(a['next_stage'], device=dev),
                    stage_of=torch.as_tensor(a['stage_of'], device=dev), state=z(4, n), out=z(4, n),
                    base=torch.as_tensor(a['stage_base'], device=dev))
        sg = _lib.StagedTasks(ms, bufs['tlist'].data_ptr(), bufs['tmeta'].data_ptr(), bufs['next_stage'].data_ptr(),
                              bufs['stage_of'].data_ptr(), int(bool(progressive)), 0, bufs['state'].data_ptr(),
                              bufs['out'].data_ptr())
        t = self._dev(a['stage_targets'], torch.int8, (ms, _lib.CELLS))
        f = self._dev(a['stage_fulls'], torch.int8, (ms, _lib.CELLS))
        self._check_block_capacity(t, None)
        with self._guard():
            _lib.check(self._L.iglu_staged_prepare(self._ctx, C.byref(sg), t.data_ptr(), f.data_ptr(), self._stream()))
            self._count()
        bufs['struct'] = sg
        self._staged = bufs
        return self.reset() if reset else None

    @property
    def staged(self):
        """Staged score of the last step (Subtasks.step_intersection on each env's grid): int32 device tensors
        right, wrong, goal (stage row; minus self._staged['base'][structure] = the turn) and bool done."""
        if self._staged is None:
            raise ValueError('no staged tasks attached (set_staged_tasks)')
        o = self._staged['out']
        return {'right': o[0], 'wrong': o[1], 'done': o[2].bool(), 'goal': o[3]}

    @property
    def task_ids(self):
        """int32 [num_envs]: pool index of each env's current task (-1 before its first draw); the host
        side of obs['dialog'] (task.chat) is looked up with it."""
        if getattr(self, '_pool', None) is None:
            raise ValueError('no task pool attached')
        return self._pool['cursor'][0]

    def _check_block_capacity(self, targets, starts):
        """A block list longer than the row capacity chosen with max_target_blocks would be truncated by the
        device-side Task.__init__; refuse instead, before any task row or pool pointer is touched (only checked
        when the capacity is below the 1089-cell maximum, because it needs a device -> host read)."""
        room = self._tlist.shape[1] - _lib.TLIST_HDR
        if room >= _lib.CELLS:
            return
        diff = targets != (starts if starts is not None else 0)   # cells of target - start (env.py:230)
        worst = int(diff.sum(1).max())
        if worst > room:
            raise ValueError(f'a task has {worst} blocks to place / remove but max_target_blocks '
                             f'allows {room}; create the env with a larger max_target_blocks')

    def _check_task(self, stepping=False):
        if not self._tasks_set:
            _lib.check(-3)  # ValueError with the reference's message (env.py:199-201)
        if stepping and not self._was_reset:
            # a never-reset env has active_block = 0 and time_int_steps = 0 (the reference sets both in
            # Agent.__init__ and always resets inside set_task): stepping it is refused like env.py:275
            raise ValueError('Task is not initialized! Run .reset() first.')

    def reset(self, mask=None, hard=False):
        """Batched GridWorld.reset (reference env.py:206-261) for envs with mask != 0 (None = all)."""
        self._check_task()
        m = None if mask is None else self._dev(mask, torch.uint8, (self.num_envs,))
        with torch.cuda.device(self.device):
            _lib.check(self._L.iglu_reset(self._ctx, C.byref(self._st), C.byref(self._ob),
                                          None if m is None else m.data_ptr(), int(hard), self._stream()))
            self._count()
        if m is None:
            self._was_reset = True
        elif not self._was_reset:  # masked reset of a fresh env: fine once every env has been covered
            self._reset_seen = m.clone() if self._reset_seen is None else (self._reset_seen | m)
            if bool(self._reset_seen.all()):
                self._was_reset, self._reset_seen = True, None
        if self.do_render:
            self.render()
        return self.observation()

    # ------------------------------------------------------------------ step
    def render(self):
        """obs['pov'] of every env for the current state (reference env.py:263-266 / render.py:128-156): uint8
        [n, h, w, 3] on the device, written by one ray-casting launch."""
        if not self.do_render:
            raise ValueError('create env with render=True')
        with self._guard():
            _lib.check(self._L.iglu_render_pov(self._ctx, C.byref(self._st), self._pov.data_ptr(), self.render_size[1],
                                               self.render_size[0], None if self._atlas is None else self._atlas.data_ptr(),
                                               self._stream()))
            self._count()
        return self._pov

    def observation(self):
        obs = {'inventory': self._obs_inv, 'compass': self._obs_compass}
        if self.do_render:
            obs['pov'] = self._pov
        if self.vector_state:
            obs['agentPos'] = self._obs_pos
            obs['grid'] = self._obs_grid if self.grid_obs else self.grid
        if self.target_in_obs:
            obs['target_grid'] = self._obs_target
        return obs

    def step(self, actions):
        """Batched GridWorld.step (reference env.py:268-303).

        walking (discretize=True): int32 [n]; flying: dict(movement f32 [n,3], camera f32 [n,2],
        inventory int [n], placement int [n]); walking (discretize=False): dict with the nine
        reference keys, each [n] (camera [n,2]).
        Returns (obs dict of device tensors, reward float64 [n], done bool [n], {}).
        """
        self._check_task(stepping=True)
        n = self.num_envs
        L, st, ob = self._L, C.byref(self._st), C.byref(self._ob)
        with self._guard():
            if self._kind == _lib.WALKING_DISCRETE:
                a = self._dev(actions, torch.int32, (n,))
                _lib.check(L.iglu_step_walking(self._ctx, st, a.data_ptr(), ob, self._stream()))
            elif self._kind == _lib.FLYING:
                mv = self._dev(actions['movement'], torch.float32, (n, 3))
                cam = self._dev(actions['camera'], torch.float32, (n, 2))
                inv = self._dev(actions['inventory'], torch.int32, (n,))
                pl = self._dev(actions['placement'], torch.int32, (n,))
                _lib.check(L.iglu_step_flying(self._ctx, st, mv.data_ptr(), cam.data_ptr(), inv.data_ptr(),
                                              pl.data_ptr(), ob, self._stream()))
            else:
                if isinstance(actions, dict):
                    keys = ('forward', 'back', 'left', 'right', 'jump', 'attack', 'use', 'hotbar')
                    b = torch.stack([self._dev(actions[k], torch.int32, (n,)) for k in keys], dim=1)
                    cam = actions['camera']
                else:
                    b, cam = actions
                b = self._dev(b, torch.int32, (n, 8))
                cam = self._dev(cam, torch.float32, (n, 2))
                _lib.check(L.iglu_step_walking_dict(self._ctx, st, b.data_ptr(), cam.data_ptr(), ob,
                                                    self._stream()))
            self._count()
            if self._staged is not None:   # Subtasks.step_intersection(env.grid) for every env, as one more launch
                _lib.check(L.iglu_staged_step(self._ctx, st, C.byref(self._staged['struct']),
                                              self._pool['cursor'].data_ptr(), self._stream()))
                self._count()
        if self.do_render:   # env.py:299-300
            self.render()
        return self.observation(), self._reward, self._done.view(torch.bool), {}

    def step_host(self, actions_host, reward_host, done_host, wait=True):
        """End-to-end host-buffer step (walking Discrete): actions int32 [n] in (pinned) host memory are
        copied to the device, the step runs, reward (float64 [n]) / done (uint8 [n]) are copied back
        into the given host tensors and the stream is synchronised. Observations stay on the device.
        wait=False only enqueues the step (iglu_step_walking_host_async): the results are valid after
        self.wait(), which lets a host loop keep a second env group in flight meanwhile."""
        if not (self._tasks_set and self._was_reset):
            self._check_task(stepping=True)
        if self._kind != _lib.WALKING_DISCRETE:
            raise NotImplementedError('step_host is provided for the Discrete(18) action space')
        n = self.num_envs
        # validated on every call (a cache keyed by address would go stale when buffers are freed and reused)
        if (actions_host.dtype != torch.int32 or reward_host.dtype != torch.float64 or done_host.dtype != torch.uint8
                or actions_host.numel() != n or reward_host.numel() != n or done_host.numel() != n
                or not (actions_host.is_contiguous() and reward_host.is_contiguous() and done_host.is_contiguous())
                or actions_host.is_cuda or reward_host.is_cuda or done_host.is_cuda):
            raise ValueError('step_host needs contiguous CPU tensors of dtype int32 / float64 / uint8, [n]')
        pa, pr, pd = actions_host.data_ptr(), reward_host.data_ptr(), done_host.data_ptr()
        fn = self._L.iglu_step_walking_host if wait else self._L.iglu_step_walking_host_async
        with self._guard():
            self._host_stream = self._stream()
            rc = fn(self._ctx, self._st_ref, pa, self._actions_ptr, self._ob_ref, pr, pd, self._host_stream)
        if rc:
            _lib.check(rc)
        self.gpu_launches += self._L.iglu_last_launch_count(self._ctx)
        return reward_host, done_host

    def wait(self):
        """Blocks until the last step_host(..., wait=False) of this env has delivered its results."""
        rc = self._L.iglu_wait(getattr(self, '_host_stream', None) or self._stream())
        if rc:
            _lib.check(rc)

    def fetch_env(self, env_idx=0):
        """One env's observation + agent state in host memory (an _lib.EnvSnapshot; one kernel, one copy, one
        synchronisation -- what the N = 1 gym surface reads after every step)."""
        if self._snap is None:
            self._snap = (torch.zeros(C.sizeof(_lib.EnvSnapshot), dtype=torch.uint8, device=self.device),
                          torch.zeros(C.sizeof(_lib.EnvSnapshot), dtype=torch.uint8).pin_memory())
            self._snap_view = _lib.EnvSnapshot.from_address(self._snap[1].data_ptr())
        with self._guard():
            _lib.check(self._L.iglu_fetch_env_host(self._ctx, self._st_ref, self._ob_ref, int(env_idx),
                                                   self._snap[0].data_ptr(), self._snap[1].data_ptr(),
                                                   self._stream()))
            self._count()
        return self._snap_view

    @property
    def evaluations(self):
        """Running count of maximal_intersection evaluations performed by step calls (device counter)."""
        return int(self._sched[2:4].view(torch.int64).item())

    def reset_evaluation_counter(self):
        self._sched[2:4] = 0

    # ------------------------------------------------------------------ state views (in place)
    @property
    def grid(self):
        """int8 [n, 9, 11, 11] view of the resident voxel state (env.grid)."""
        return self._grid[:, :_lib.CELLS].view(self.num_envs, 9, 11, 11)

    @property
    def grid_rows(self):
        """int8 [n, 1104]: the resident voxel state as it lies in HBM -- 1089 cells + zero pad per env, rows 16-byte
        aligned. What the first matmul of an on-device policy wants (K = 1104; no gather of the [n, 9, 11, 11] view)."""
        return self._grid

    @property
    def target_grid(self):
        return self._target[:, :_lib.CELLS].view(self.num_envs, 9, 11, 11)

    @property
    def starting_grid(self):
        return self._init[:, :_lib.CELLS].view(self.num_envs, 9, 11, 11)

    @property
    def position(self):
        return self._pose[0:3].t()

    @property
    def rotation(self):
        """(yaw, pitch) per env, fp64."""
        return self._pose[3:5].t()

    @property
    def dy(self):
        return self._pose[5]

    @property
    def inventory(self):
        w = self._istate[0:3].t().contiguous().view(torch.int16)
        return w.view(self.num_envs, 6).to(torch.int32)

    @property
    def active_block(self):
        return self._istate[3] & 0xFF

    @property
    def time_int_steps(self):
        return (self._istate[3] >> 8) & 0xFF

    @property
    def error_flags(self):
        """1 where the reference would have raised ValueError('Bad inventory id') (world.py:354-355)."""
        return (self._istate[3] >> 24) & 0xFF

    @property
    def step_no(self):
        return self._istate[4]

    @property
    def max_int(self):
        """_synthetic_task.max_int per env."""
        return self._istate[6]

    @property
    def target_size(self):
        return self._tmeta[1]

    def set_agent_state(self, env_idx, position=None, rotation=None, dy=None, time_int_steps=None,
                        active_block=None):
        """Overwrites carried agent state (tests / checkpoint restore)."""
        i = int(env_idx)
        if position is not None:
            self._pose[0:3, i] = torch.as_tensor(position, dtype=torch.float64)
        if rotation is not None:
            self._pose[3:5, i] = torch.as_tensor(rotation, dtype=torch.float64)
        if dy is not None:
            self._pose[5, i] = float(dy)
        w3 = int(self._istate[3, i])
        if time_int_steps is not None:
            w3 = (w3 & ~0xFF00) | (int(time_int_steps) << 8)
        if active_block is not None:
            w3 = (w3 & ~0xFF) | int(active_block)
        self._istate[3, i] = w3

    def set_grids(self, grids):
        """Installs pre-built worlds: env e's voxel grid becomes grids[e] (int8 [n, 9, 11, 11]) as if its agent
        had already placed those blocks -- inventories (20 per colour minus the blocks present, env.py:243-246
        and world.py:323), the block count and the cached max_int of the synthetic task (task.py:103-119) follow,
        so stepping continues exactly as the reference would from that world state."""
        n = self.num_envs
        g = self._dev(grids, torch.int8, (n, _lib.CELLS))
        init, target = self._init[:, :_lib.CELLS], self._target[:, :_lib.CELLS]
        self._grid[:, :_lib.CELLS] = g
        counts = torch.stack([(g == c).sum(1) for c in range(1, 7)], dim=1)
        inv = (20 - counts).to(torch.int16).contiguous().view(torch.int32)       # [n, 3] packed pairs
        self._istate[0:3] = inv.t()
        self._istate[5] = (g != init).sum(1).to(torch.int32)
        self._istate[6] = max_intersection(g - init, target - init, device=self.device)
        self._obs_inv.copy_((20 - counts).to(torch.float32))
        if self._obs_grid is not None:
            self._obs_grid.view(n, _lib.CELLS).copy_(g)

    def state_dict(self):
        keys = ('_grid', '_init', '_target', '_pose', '_init_pose', '_istate', '_tlist', '_tmeta')
        sd = {k[1:]: getattr(self, k).clone() for k in keys}
        if self._pool is not None:  # the generator's per-env position (the pool itself is set_task_pool's input)
            sd['pool_cursor'] = self._pool['cursor'].clone()
        sd['was_reset'] = torch.tensor(int(self._was_reset))
        return sd

    def load_state_dict(self, sd):
        for k, v in sd.items():
            if k == 'pool_cursor':
                if self._pool is None:
                    raise ValueError('state has a task-pool cursor: call set_task_pool(..., reset=False) first')
                self._pool['cursor'].copy_(v)
            elif k == 'was_reset':
                self._was_reset = bool(int(v))
            else:
                getattr(self, '_' + k).copy_(v)
        self._tasks_set = True
        if 'was_reset' not in sd:
            self._was_reset = True
        # observation buffers the step path does not rewrite in full are derived state: rebuild them
        if self._obs_grid is not None:
            self._obs_grid.view(self.num_envs, _lib.CELLS).copy_(self._grid[:, :_lib.CELLS])
        if self._obs_target is not None:
            self._obs_target.view(self.num_envs, _lib.CELLS).copy_(self._target[:, :_lib.CELLS])


def load_texture(path='auto'):
    """The 128 x 128 RGBA texture atlas as a uint8 array [128, 128, 4] (rows top-down), or None.

    path: a PNG file, or 'auto' = the `texture.png` that ships inside an installed reference package
    (`gridworld/texture.png`, reference render.py:54, 64-67) if that package can be located -- it is NOT imported --
    else None (the ray caster then uses flat colours). Needs Pillow; returns None without it."""
    import importlib.util
    import os
    if path == 'auto':
        try:
            spec = importlib.util.find_spec('gridworld')
        except (ImportError, ValueError):
            spec = None
        roots = list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []
        path = next((os.path.join(r, 'texture.png') for r in roots if os.path.exists(os.path.join(r, 'texture.png'))), None)
        if path is None:
            return None
    try:
        from PIL import Image
    except ImportError:
        return None
    a = np.asarray(Image.open(path).convert('RGBA'), dtype=np.uint8)
    if a.shape != (128, 128, 4):
        raise ValueError(f'{path}: expected a 128 x 128 texture atlas, got {a.shape}')
    return a


def make_vec(num_envs, device='cuda', **kwargs):
    """Batched counterpart of gym.make('IGLUGridworld-v0', ...) (reference env.py:333-362)."""
    kwargs.pop('size_reward', None)
    kwargs.pop('name', None)
    kwargs.pop('fake', None)
    return IGLUVecEnv(num_envs, device=device, **kwargs)


def sample_tasks(m, max_blocks=4, height_levels=1, max_dist=2, num_colors=1, seed=0, first_task=0, device='cuda'):
    """m random target structures with the semantics of RandomTasks.sample_task (reference task_set.py:135-157),
    generated on the GPU by iglu_sample_tasks; int8 device tensor [m, 9, 11, 11]. The random stream is the
    counter-based one documented in include/iglu_b200.h (task `first_task + i`, so shards can split a pool)."""
    dev = _require_cuda(device)
    L = _lib.load()
    out = torch.empty(int(m), 9, 11, 11, dtype=torch.int8, device=dev)
    spec = _lib.RandomTaskSpec(int(max_blocks), int(height_levels), int(max_dist), int(num_colors))
    with torch.cuda.device(dev):
        _lib.check(L.iglu_sample_tasks(out.data_ptr(), int(m), C.byref(spec), int(seed) & (2 ** 64 - 1),
                                       int(first_task), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return out


def max_intersection(grids, targets, full_grids=None, invariant=True, return_argmax=False, device='cuda'):
    """Task(target, full_grid=full, invariant=invariant).maximal_intersection(grid) for a batch of pairs
    (reference gridworld/tasks/task.py:9-72, 147-161), on the GPU. Inputs: int8-convertible
    [k, 9, 11, 11]. Returns int32 [k] (and int32 [k, 3] = (dx, dz, rot) like argmax_intersection).
    Device int8 tensors of shape [k, 1104] -- rows in the env state layout, e.g. IGLUVecEnv._grid / ._target --
    are evaluated where they lie (iglu_max_intersection_strided), without a repack."""
    dev = _require_cuda(device)
    L = _lib.load()
    g = torch.as_tensor(grids)
    t = torch.as_tensor(targets)
    rows = lambda x: (x.dim() == 2 and x.shape[1] == _lib.GRID_STRIDE and x.dtype == torch.int8  # noqa: E731
                      and x.is_cuda and x.is_contiguous())
    if rows(g) and rows(t) and full_grids is None:
        k = g.shape[0]
        out = torch.zeros(k, dtype=torch.int32, device=g.device)
        arg = torch.zeros(k, 3, dtype=torch.int32, device=g.device) if return_argmax else None
        with torch.cuda.device(g.device):
            _lib.check(L.iglu_max_intersection_strided(
                g.data_ptr(), _lib.GRID_STRIDE, t.data_ptr(), _lib.GRID_STRIDE, None, 0, int(bool(invariant)),
                out.data_ptr(), None if arg is None else arg.data_ptr(), k,
                C.c_void_p(torch.cuda.current_stream(g.device).cuda_stream)))
        return (out, arg) if return_argmax else out
    k = g.shape[0] if g.dim() in (2, 4) else 1   # [k, 9, 11, 11] or [k, 1089]; a single grid otherwise

    def prep(x):
        if x is None:
            return None
        return torch.as_tensor(x).to(device=dev, dtype=torch.int8).reshape(k, _lib.CELLS).contiguous()

    g, t, f = prep(grids), prep(targets), prep(full_grids)
    out = torch.zeros(k, dtype=torch.int32, device=dev)
    arg = torch.zeros(k, 3, dtype=torch.int32, device=dev) if return_argmax else None
    with torch.cuda.device(dev):
        _lib.check(L.iglu_max_intersection(g.data_ptr(), t.data_ptr(), None if f is None else f.data_ptr(),
                                           int(bool(invariant)), out.data_ptr(),
                                           None if arg is None else arg.data_ptr(), k,
                                           C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return (out, arg) if return_argmax else out
