"""Import the UNMODIFIED Python reference from /root/reference (this container only).

TEST INFRASTRUCTURE ONLY (used by oracle/make_golden.py and, when /root/reference is
present, by tests that cross-check the C oracle against the live reference).
Two shims, both recorded in SURVEY.md 8(c):
  1. `gym` stand-in (oracle/refharness/gym) because gym/gymnasium are not installed;
  2. `numpy.int = int` because reference gridworld/tasks/task.py:171 uses the removed alias.
"""
import os
import sys

REFERENCE_ROOT = os.environ.get('IGLU_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'gridworld'))


def load():
    """Returns the imported reference `gridworld` package (and registers its envs in the gym stub)."""
    if not available():
        raise RuntimeError(f'reference tree not found at {REFERENCE_ROOT}')
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)          # makes `import gym` resolve to the stub
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import numpy as np
    if not hasattr(np, 'int'):
        np.int = int
    import warnings
    warnings.filterwarnings('ignore')
    import gridworld  # noqa: F401  (the reference; registers IGLUGridworld-v0)
    return gridworld
