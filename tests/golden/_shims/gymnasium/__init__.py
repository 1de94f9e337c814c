"""Test-only stand-in for `gymnasium`, used ONLY by tests/golden/make_golden.py so the
reference's frozenlake/FrozenLakeGame.py (which needs gymnasium solely for the map
description, FrozenLakeGame.py:26-43) imports unmodified in the build container."""
import numpy as np

_MAPS = {
    "FrozenLake-v1": ["SFFF", "FHFH", "FFFH", "HFFG"],
    "FrozenLake8x8-v1": ["SFFFFFFF", "FFFFFFFF", "FFFHFFFF", "FFFFFHFF",
                         "FFFHFFFF", "FHHFFFHF", "FHFFHFHF", "FFFHFFFG"],
}


class _Env:
    def __init__(self, desc):
        self.desc = np.asarray([[c.encode() for c in row] for row in desc], dtype="|S1")
        self.unwrapped = self

    def reset(self, **kw):
        return 0, {}

    def render(self):
        return None


def make(name, desc=None, is_slippery=False, render_mode=None, **kw):
    return _Env(desc if desc is not None else _MAPS[name])
