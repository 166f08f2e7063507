"""Stub of gym.envs: register / make (see gym/__init__.py docstring)."""
import importlib

registry = {}


def register(id, entry_point, kwargs=None, **_):
    registry[id] = (entry_point, dict(kwargs or {}))


def make(id, **kwargs):
    entry_point, kw = registry[id]
    kw = dict(kw, **kwargs)
    if callable(entry_point):
        return entry_point(**kw)
    mod, fn = entry_point.split(':')
    return getattr(importlib.import_module(mod), fn)(**kw)
