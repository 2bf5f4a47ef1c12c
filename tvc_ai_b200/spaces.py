"""Gymnasium-compatible spaces.  Uses gymnasium.spaces.Box when gymnasium is importable (it is not
in this image), otherwise a duck-typed Box with the attributes the reference trainer touches
(scripts/train.py:335-342, :562): shape, low, high, dtype, sample(), seed(), contains()."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - gymnasium is absent in the build image
    import gymnasium as _gym
    from gymnasium.spaces import Box as _GymBox
    HAVE_GYMNASIUM = True
except Exception:  # noqa: BLE001
    _gym = None
    _GymBox = None
    HAVE_GYMNASIUM = False


class _Box:
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        if shape is None:
            shape = np.shape(low)
        self.dtype = np.dtype(dtype)
        self.low = np.broadcast_to(np.asarray(low, self.dtype), shape).copy()
        self.high = np.broadcast_to(np.asarray(high, self.dtype), shape).copy()
        self.shape = tuple(shape)
        self._np_random = np.random.default_rng(seed)

    def seed(self, seed=None):
        self._np_random = np.random.default_rng(seed)
        return [seed]

    def sample(self):
        return self._np_random.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    __contains__ = contains

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype.name})"

    def __eq__(self, other):
        return (isinstance(other, _Box) and self.shape == other.shape and np.array_equal(self.low, other.low)
                and np.array_equal(self.high, other.high))


Box = _GymBox if HAVE_GYMNASIUM else _Box
EnvBase = _gym.Env if HAVE_GYMNASIUM else object

# enhanced_rocket_tvc_env.py:358-379 (declarative bounds, quirk Q21: never enforced)
OBS_LOW = np.array([-1, -1, -1, -1, -10, -10, -10, 0, 0, 0], np.float32)
OBS_HIGH = np.array([1, 1, 1, 1, 10, 10, 10, 1, 1, 1], np.float32)


def observation_space():
    return Box(low=OBS_LOW, high=OBS_HIGH, dtype=np.float32)


def action_space():
    return Box(low=-1.0, high=1.0, shape=(2,), dtype=np.float32)


def batch_space(space, n):
    low = np.broadcast_to(space.low, (n,) + tuple(space.shape)).copy()
    high = np.broadcast_to(space.high, (n,) + tuple(space.shape)).copy()
    return Box(low=low, high=high, dtype=np.float32)
