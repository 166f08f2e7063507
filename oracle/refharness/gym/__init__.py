"""Minimal stand-in for the `gym` package (gym is not installed in this image).

TEST INFRASTRUCTURE ONLY. It exposes exactly the names the reference imports
(reference gridworld/env.py:6-8,352-362 and gridworld/wrappers.py:5-6) so that the
UNMODIFIED reference can be imported from /root/reference by oracle/make_golden.py
to generate the golden vectors under tests/golden/. Nothing in the product path
imports this.
"""
from . import spaces, envs  # noqa: F401
from .envs import register, make  # noqa: F401


class Env:
    action_space = None
    observation_space = None

    @property
    def unwrapped(self):
        return self

    def reset(self):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self.action_space = env.action_space
        self.observation_space = env.observation_space

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def reset(self, **kw):
        return self.env.reset(**kw)

    def step(self, action):
        return self.env.step(action)


class ObservationWrapper(Wrapper):
    def reset(self, **kw):
        return self.observation(self.env.reset(**kw))

    def step(self, action):
        o, r, d, i = self.env.step(action)
        return self.observation(o), r, d, i


class ActionWrapper(Wrapper):
    def step(self, action):
        return self.env.step(self.action(action))
