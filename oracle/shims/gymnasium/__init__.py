"""Minimal stand-in for gymnasium 0.29.0 (see ../README.md): Env seeding contract, spaces, registry."""
import importlib

import numpy as np

from . import spaces  # noqa: F401
from .envs import registration as _reg
from .utils import seeding


class Env:
    metadata = {"render_modes": []}
    render_mode = None
    _np_random = None

    def reset(self, *, seed=None, options=None):
        # gymnasium/core.py (0.29): only re-seed when a seed is passed; otherwise the stream continues
        if seed is not None:
            self._np_random, seed = seeding.np_random(seed)

    @property
    def np_random(self) -> np.random.Generator:
        if self._np_random is None:
            self._np_random, _ = seeding.np_random()
        return self._np_random

    @np_random.setter
    def np_random(self, value):
        self._np_random = value

    def close(self):
        pass


def make(id, **kwargs):
    entry = _reg.registry[id]
    mod, cls = entry.split(":")
    return getattr(importlib.import_module(mod), cls)(**kwargs)
