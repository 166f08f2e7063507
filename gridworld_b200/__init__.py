"""gridworld_b200 -- B200-native batched implementation of the IGLU Gridworld step hot path.

Drop-in surface (reference gridworld/env.py): make('IGLUGridworld-v0', action_space='walking'|'flying',
render=False, vector_state=True, ...), env.set_task / reset / step. Batched surface: make_vec(num_envs, ...).
"""
from .tasks import Task, Tasks, enumerate_tasks, load_goals, load_tasks_dataset, pool_arrays  # noqa: F401
from .env import Actions, GridWorld, SizeReward, Wrapper, create_env, make  # noqa: F401
from .vec_env import IGLUVecEnv, load_texture, make_vec, max_intersection, sample_tasks  # noqa: F401

__all__ = ['Task', 'Tasks', 'enumerate_tasks', 'load_goals', 'load_tasks_dataset', 'pool_arrays', 'Actions', 'GridWorld', 'SizeReward',
           'Wrapper', 'create_env', 'make', 'IGLUVecEnv', 'make_vec', 'max_intersection', 'sample_tasks', 'load_texture']
