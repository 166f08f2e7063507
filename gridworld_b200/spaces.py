"""Minimal gym-compatible spaces (gym / gymnasium are not installed in this image).

Mirrors what the reference builds in GridWorld.__init__ (reference gridworld/env.py:12-21, 56-95).
If `gym` is importable its classes are used instead so isinstance checks in user code keep working.
"""
import numpy as np

try:  # pragma: no cover - gym is absent in the build image
    from gym.spaces import Space, Box, Discrete, Dict  # type: ignore
    HAVE_GYM = True
except Exception:  # noqa: BLE001
    HAVE_GYM = False

    class Space:
        def __init__(self, shape=None, dtype=None):
            self.shape = None if shape is None else tuple(shape)
            self.dtype = None if dtype is None else np.dtype(dtype)

        def sample(self):
            raise NotImplementedError

        def contains(self, x):
            raise NotImplementedError

        def __contains__(self, x):
            return self.contains(x)

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            if shape is None:
                shape = np.asarray(low).shape
            super().__init__(shape, dtype)
            self.low = np.broadcast_to(np.asarray(low, dtype=dtype), shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=dtype), shape).copy()

        def sample(self):
            if np.issubdtype(self.dtype, np.integer):
                return np.random.randint(self.low, self.high + 1).astype(self.dtype)
            return np.random.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f'Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})'

    class Discrete(Space):
        def __init__(self, n):
            super().__init__((), np.int64)
            self.n = int(n)

        def sample(self):
            return int(np.random.randint(self.n))

        def contains(self, x):
            try:
                return 0 <= int(x) < self.n
            except (TypeError, ValueError):
                return False

        def __repr__(self):
            return f'Discrete({self.n})'

    class Dict(Space):
        def __init__(self, spaces):
            super().__init__(None, None)
            self.spaces = dict(spaces)

        def __getitem__(self, k):
            return self.spaces[k]

        def keys(self):
            return self.spaces.keys()

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}

        def contains(self, x):
            return isinstance(x, dict) and all(k in x and s.contains(x[k]) for k, s in self.spaces.items())

        def __repr__(self):
            return 'Dict(' + ', '.join(f'{k}: {v!r}' for k, v in self.spaces.items()) + ')'


class String(Space):
    """reference gridworld/env.py:12-21"""

    def __init__(self):
        super().__init__(shape=(), dtype=np.object_)

    def sample(self):
        return ''

    def contains(self, obj):
        return isinstance(obj, str)
