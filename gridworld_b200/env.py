"""Drop-in single-env gym surface over the batched CUDA path (N = 1).

Mirrors reference gridworld/env.py: GridWorld (26-303), Wrapper (306-314), SizeReward (316-331),
create_env (333-350) and the 'IGLUGridworld-v0' / 'IGLUGridworldVector-v0' registrations (352-362).
Observations come back as numpy arrays with the reference's keys, dtypes and shapes; all state
updates run in libiglu_b200.so on the GPU (no CPU fallback). render=True adds obs['pov'] from a CUDA voxel ray caster.
"""
import warnings

import numpy as np

from .tasks import Task, Tasks
from .vec_env import IGLUVecEnv


class _AgentView:
    """Read/write view of the device-resident Agent (reference core/world.py:8-29). Reads come from the env's
    last snapshot (one device -> host transfer per step, GridWorld._snapshot), writes go to the device."""
    PAD = 0.25

    def __init__(self, owner, flying):
        self._o = owner
        self.flying = flying
        self.sustain = False
        self.strafe = [0, 0]

    def _write(self, **kw):
        self._o._vec.set_agent_state(0, **kw)
        self._o._snap = None  # stale now

    @property
    def position(self):
        return tuple(self._o._snapshot().pose[0:3])

    @position.setter
    def position(self, p):
        self._write(position=p)

    @property
    def rotation(self):
        return tuple(self._o._snapshot().pose[3:5])

    @rotation.setter
    def rotation(self, r):
        self._write(rotation=r)

    @property
    def dy(self):
        return float(self._o._snapshot().pose[5])

    @dy.setter
    def dy(self, v):
        self._write(dy=v)

    @property
    def time_int_steps(self):
        return int(self._o._snapshot().time_int_steps)

    @property
    def active_block(self):
        return int(self._o._snapshot().active_block)

    @property
    def inventory(self):
        return list(self._o._snapshot().inventory)


class GridWorld:
    """reference gridworld/env.py:26-303 with render=False."""

    def __init__(self, render=True, max_steps=250, select_and_place=False, discretize=False,
                 right_placement_scale=1., wrong_placement_scale=0.1, render_size=(64, 64),
                 target_in_obs=False, action_space='walking', vector_state=True, fake=False, name='',
                 device='cuda', texture='auto'):
        # render=True: obs['pov'] comes from the CUDA voxel ray caster (csrc/iglu_render.cuh), the stand-in for the
        # reference's pyglet / OpenGL renderer; `texture` = the reference's texture atlas (array / PNG path / 'auto' = the
        # texture.png of an installed reference package when there is one; None = flat colours)
        self._vec = IGLUVecEnv(1, device=device, action_space=action_space, discretize=discretize,
                               render=bool(render) and not fake, render_size=render_size, texture=texture,
                               select_and_place=select_and_place, max_steps=max_steps,
                               right_placement_scale=right_placement_scale,
                               wrong_placement_scale=wrong_placement_scale, target_in_obs=target_in_obs,
                               vector_state=vector_state, auto_reset=False, reward_mode='on_change',
                               grid_obs=True)
        self._snap = None
        self.agent = _AgentView(self, action_space == 'flying')
        self.world = None
        self._task = None
        self._task_generator = None
        self.step_no = 0
        self.right_placement_scale = right_placement_scale
        self.wrong_placement_scale = wrong_placement_scale
        self.max_steps = max_steps
        self.right_placement = 0
        self.wrong_placement = 0
        self.render_size = render_size
        self.select_and_place = select_and_place
        self.target_in_obs = target_in_obs
        self.vector_state = vector_state
        self.discretize = discretize
        self.action_space_type = action_space
        self.starting_grid = None
        self.fake = fake
        self._overwrite_starting_grid = None
        self.initial_position = (0, 0, 0)
        self.initial_rotation = (0, 0)
        self.action_space = self._vec.action_space
        self.observation_space = self._vec.observation_space
        if render and fake:   # env.py:93-94: the space has 'pov' even when the frames are fake samples
            from .vec_env import make_observation_space
            self.observation_space = make_observation_space(vector_state, target_in_obs, True,
                                                            (int(render_size[1]), int(render_size[0])))
        self.max_int = 0
        self.name = name
        self.do_render = bool(render)
        self.renderer = None

    @property
    def unwrapped(self):
        return self

    def _snapshot(self):
        """Host copy of this env's observation + agent state (iglu_fetch_env_host: one kernel, one copy, one
        synchronisation), cached until the device state changes again."""
        if self._snap is None:
            self._snap = self._vec.fetch_env(0)
        return self._snap

    @property
    def grid(self):
        return np.ctypeslib.as_array(self._snapshot().grid).reshape(9, 11, 11).copy()

    def set_task(self, task):
        """env.py:155-166"""
        if self._task_generator is not None:
            warnings.warn("The .set_task method has no effect with an initialized tasks generator. "
                          "Drop it using .set_tasks_generator(None) after calling .set_task")
        self._task = task
        self.reset()

    def set_task_generator(self, task_generator):
        """env.py:168-175"""
        self._task_generator = task_generator
        self.reset()

    def initialize_world(self, starting_grid, initial_poisition):
        """env.py:177-187"""
        self._overwrite_starting_grid = starting_grid
        warnings.warn('Default task starting grid is overwritten using .initialize_world method. '
                      'Use .deinitialize_world to restore the original state.')
        self.initial_position = tuple(initial_poisition[:3])
        self.initial_rotation = tuple(initial_poisition[3:])
        self.reset()

    def deinitialize_world(self):
        """env.py:189-193"""
        self._overwrite_starting_grid = None
        self.initial_position = (0, 0, 0)
        self.initial_rotation = (0, 0)
        self.reset()

    _NO_TASK = ('Task is not initialized! Initialize task before working with'
                ' the environment using .set_task method OR set tasks distribution using '
                '.set_task_generator method')

    @property
    def task(self):
        """env.py:195-204"""
        if self._task is None:
            if self._task_generator is None:
                raise ValueError(self._NO_TASK)
            self._task = self._task_generator.reset()
            self.starting_grid = self._task.starting_grid
        return self._task

    def _obs_numpy(self):
        sn = self._snapshot()
        as_np = np.ctypeslib.as_array
        obs = {'inventory': as_np(sn.inventory_obs).copy(), 'compass': np.array([sn.compass], dtype=np.float32),
               'dialog': self._task.chat}
        if self.vector_state:
            obs['grid'] = as_np(sn.grid).reshape(9, 11, 11).copy()
            obs['agentPos'] = as_np(sn.agent_pos).copy()
        if self.target_in_obs:
            obs['target_grid'] = as_np(sn.target_grid).reshape(9, 11, 11).copy()
        if self.do_render and not self.fake:      # env.py:257-258, 299-300
            obs['pov'] = self._vec._pov[0].cpu().numpy()
        elif self.do_render:                      # env.py:259-260, 301-302
            obs['pov'] = self.observation_space['pov'].sample()
        return obs

    def reset(self):
        """env.py:206-261"""
        if self._task is None:
            if self._task_generator is None:
                raise ValueError(self._NO_TASK)
            self._task = self._task_generator.reset()
        elif self._task_generator is not None:
            self._task = self._task_generator.reset()
        self.step_no = 0
        self._task.reset()
        if self._overwrite_starting_grid is not None:
            self.starting_grid = self._overwrite_starting_grid
        else:
            self.starting_grid = self._task.starting_grid
        # The reference crashes in step() when starting_grid is None (env.py:224-232, 290); here None is
        # treated as the empty list [] (the only way the reference's own README loop can run).
        start = self.starting_grid if self.starting_grid is not None else []
        start_dense = np.asarray(Tasks.to_dense(start)).astype(np.int8)
        target = np.asarray(self._task.target_grid).astype(np.int8)
        pose = np.array([[*self.initial_position, *self.initial_rotation]], dtype=np.float64)
        self._vec.set_tasks(target[None], start_dense[None], pose, reset=True)
        self._snap = None
        grid = self.grid
        self.max_int = self._task.maximal_intersection(grid)   # env.py:241 (never refreshed by step)
        self.prev_grid_size = int(np.count_nonzero(grid))
        return self._obs_numpy()

    def render(self):
        """env.py:263-266; RGB frame, uint8 [h, w, 3] (the reference returns RGBA and drops alpha for obs['pov'])"""
        if not self.do_render or self.fake:
            raise ValueError('create env with render=True')
        return self._vec.render()[0].cpu().numpy()

    def step(self, action):
        """env.py:268-303"""
        if self._task is None:
            if self._task_generator is None:
                raise ValueError(self._NO_TASK)
            raise ValueError('Task is not initialized! Run .reset() first.')
        v = self._vec
        if self.action_space_type == 'flying':
            if action['inventory'] < 0 or action['inventory'] > 6:
                raise ValueError(f"Bad inventory id: {action['inventory']}")  # world.py:354-355
            a = {'movement': np.asarray(action['movement'], np.float32)[None],
                 'camera': np.asarray(action['camera'], np.float32)[None],
                 'inventory': np.asarray([action['inventory']], np.int32),
                 'placement': np.asarray([action['placement']], np.int32)}
        elif self.discretize:
            a = np.asarray([int(action)], np.int32)
        else:
            if action['hotbar'] < 0 or action['hotbar'] > 6:
                raise ValueError(f"Bad inventory id: {action['hotbar']}")
            a = {k: np.asarray([int(action[k])], np.int32)
                 for k in ('forward', 'back', 'left', 'right', 'jump', 'attack', 'use', 'hotbar')}
            a['camera'] = np.asarray(action['camera'], np.float32)[None]
        prev_max = self._snapshot().max_int
        v.step(a)
        self._snap = None
        sn = self._snapshot()      # the step's single device -> host transfer
        self.step_no = int(sn.step_no)
        obs = self._obs_numpy()
        r = float(sn.reward)
        if sn.max_int != prev_max and isinstance(self.right_placement_scale, int):
            r = int(r)  # right_placement * int scale is a Python int in the reference (env.py:296)
        return obs, r, bool(sn.done), {}


class Wrapper:
    """reference env.py:306-314 (gym.Wrapper with attribute passthrough)"""

    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name == 'env':
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def reset(self):
        return self.env.reset()

    def step(self, action):
        return self.env.step(action)

    def render(self, mode='human', **kwargs):
        return self.env.render()


class Actions(Wrapper):
    """reference gridworld/wrappers.py:11-32: Discrete(17) action space without 'place' (hotbar places)."""

    def __init__(self, env):
        super().__init__(env)
        from .spaces import Discrete
        self.action_map = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16]
        self.action_space = Discrete(len(self.action_map))

    def step(self, action):
        return self.env.step(self.action_map[action])


class SizeReward(Wrapper):
    """reference env.py:316-331. NB: in the reference GridWorld.step never updates .max_int /
    .wrong_placement, so this wrapper yields max_int-at-reset on the first step and 0 afterwards
    (SURVEY section 7); that behaviour is reproduced, not fixed."""

    def __init__(self, env):
        super().__init__(env)
        self.size = 0

    def reset(self):
        self.size = 0
        return super().reset()

    def step(self, action):
        obs, reward, done, info = super().step(action)
        intersection = self.unwrapped.max_int
        reward = max(intersection, self.size) - self.size
        self.size = max(intersection, self.size)
        reward += min(self.unwrapped.wrong_placement * 0.02, 0)
        return obs, reward, done, info


def create_env(render=True, discretize=True, size_reward=True, select_and_place=True, right_placement_scale=1,
               render_size=(64, 64), target_in_obs=False, vector_state=False, max_steps=250,
               action_space='walking', wrong_placement_scale=0.1, name='', fake=False, device='cuda'):
    """reference env.py:333-350"""
    env = GridWorld(render=render, select_and_place=select_and_place, discretize=discretize,
                    right_placement_scale=right_placement_scale, wrong_placement_scale=wrong_placement_scale,
                    name=name, render_size=render_size, target_in_obs=target_in_obs, vector_state=vector_state,
                    max_steps=max_steps, action_space=action_space, fake=fake, device=device)
    if size_reward:
        env = SizeReward(env)
    return env


# ---- registry (reference env.py:352-362); also registered in gym when gym is importable ----
_REGISTRY = {
    'IGLUGridworld-v0': {},
    'IGLUGridworldVector-v0': {'vector_state': True, 'render': False},
}


def make(id, **kwargs):
    if id not in _REGISTRY:
        raise KeyError(f'unknown env id {id!r}; known: {sorted(_REGISTRY)}')
    return create_env(**dict(_REGISTRY[id], **kwargs))


try:  # pragma: no cover - gym is absent in the build image
    import gym as _gym
    for _id, _kw in _REGISTRY.items():
        try:
            _gym.envs.register(id=_id, entry_point='gridworld_b200.env:create_env', kwargs=_kw)
        except Exception:  # noqa: BLE001  (already registered)
            pass
except Exception:  # noqa: BLE001
    pass
