"""Minimal stand-in for the `gymnasium` package -- TEST INFRASTRUCTURE ONLY.

The reference env (envs/manipulation_env.py:9-10,14,85-106,140-143) uses exactly
three things from gymnasium, and gymnasium is not installed in this image:

  * ``gym.Env`` as a base class whose ``reset(seed=...)`` (re)seeds ``self.np_random``
  * the lazily created ``self.np_random`` generator
  * ``spaces.Box(low, high, shape, dtype)`` with ``.low/.high/.shape/.dtype/.sample()``

This module provides those with the semantics of gymnasium >= 0.29
(``gymnasium.utils.seeding.np_random``: ``Generator(PCG64(SeedSequence(seed)))``).
It is put on ``sys.path`` only by ``oracle/ref_harness.py`` (the importer of the real
reference), by the golden-vector generator and by ``bench.py --impl reference``.
The product package never imports it.
"""
import numpy as np

from . import spaces  # noqa: F401  (``from gymnasium import spaces``)

__version__ = "0.29.stub"


def _make_rng(seed=None):
    return np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))


class Env:
    metadata = {"render_modes": []}
    render_mode = None
    spec = None
    _np_random = None

    def reset(self, *, seed=None, options=None):
        # gymnasium only re-seeds when a seed is given
        if seed is not None:
            self._np_random = _make_rng(seed)

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random = _make_rng(None)
        return self._np_random

    @np_random.setter
    def np_random(self, value):
        self._np_random = value

    @property
    def unwrapped(self):
        return self

    def step(self, action):
        raise NotImplementedError

    def render(self):
        raise NotImplementedError

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *args):
        self.close()
        return False
