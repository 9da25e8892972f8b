"""Minimal stand-in for the slice of `gym` 0.20 the maze API touches.

Used only when `import gym` fails (gym is not installed in the build image).
Covers what the reference itself uses — gym.Env (maze_env.py:27), spaces.Box
(maze_env.py:246, point.py:41), envs.register (__init__.py:22) — and what its
tests / README use: gym.make, env.action_space.sample(), env.unwrapped, and the
TimeLimit wrapper that `max_episode_steps=1000` implies (__init__.py:31).
"""

import importlib
import types
from typing import Any, Callable, Dict, Optional

import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)
        self._rng = np.random.default_rng()

    def seed(self, seed: Optional[int] = None):
        self._rng = np.random.default_rng(seed)
        return [seed]

    def sample(self):
        raise NotImplementedError

    def contains(self, x) -> bool:
        raise NotImplementedError

    def __contains__(self, x) -> bool:
        return self.contains(x)


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        low, high = np.asarray(low, dtype=dtype), np.asarray(high, dtype=dtype)
        if shape is not None:
            low, high = np.broadcast_to(low, shape).copy(), np.broadcast_to(high, shape).copy()
        assert low.shape == high.shape
        super().__init__(low.shape, dtype)
        self.low, self.high = low, high
        self.bounded_below = np.isfinite(low)
        self.bounded_above = np.isfinite(high)

    def sample(self) -> np.ndarray:
        lo_b, hi_b = self.bounded_below, self.bounded_above
        out = np.empty(self.shape, dtype=np.float64)
        both, neither = lo_b & hi_b, ~lo_b & ~hi_b
        only_lo, only_hi = lo_b & ~hi_b, ~lo_b & hi_b
        out[both] = self._rng.uniform(self.low[both], self.high[both])
        out[neither] = self._rng.normal(size=int(neither.sum()))
        out[only_lo] = self.low[only_lo] + self._rng.exponential(size=int(only_lo.sum()))
        out[only_hi] = self.high[only_hi] - self._rng.exponential(size=int(only_hi.sum()))
        return out.astype(self.dtype)

    def contains(self, x) -> bool:
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    def __eq__(self, other):
        return isinstance(other, Box) and self.shape == other.shape and \
            np.allclose(self.low, other.low) and np.allclose(self.high, other.high)


class Env:
    metadata: Dict[str, Any] = {"render.modes": []}
    reward_range = (-float("inf"), float("inf"))
    spec = None
    action_space: Space = None
    observation_space: Space = None

    def step(self, action):
        raise NotImplementedError

    def reset(self, **kwargs):
        raise NotImplementedError

    def render(self, mode="human", **kwargs):
        raise NotImplementedError

    def close(self):
        pass

    def seed(self, seed=None):
        return [seed]

    @property
    def unwrapped(self):
        return self

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


class Wrapper(Env):
    def __init__(self, env: Env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def action_space(self):
        return self.env.action_space

    @property
    def observation_space(self):
        return self.env.observation_space

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def render(self, mode="human", **kwargs):
        return self.env.render(mode=mode, **kwargs)

    def close(self):
        return self.env.close()


class TimeLimit(Wrapper):
    """gym 0.20 semantics: done=True once max_episode_steps elapsed; info flags truncation."""

    def __init__(self, env: Env, max_episode_steps: Optional[int] = None):
        super().__init__(env)
        self._max_episode_steps = max_episode_steps
        self._elapsed_steps = None

    def step(self, action):
        assert self._elapsed_steps is not None, "Cannot call env.step() before calling reset()"
        obs, reward, done, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._elapsed_steps >= self._max_episode_steps:
            info["TimeLimit.truncated"] = not done
            done = True
        return obs, reward, done, info

    def reset(self, **kwargs):
        self._elapsed_steps = 0
        return self.env.reset(**kwargs)


class EnvSpec:
    def __init__(self, id, entry_point, kwargs=None, max_episode_steps=None, reward_threshold=None):
        self.id, self.entry_point = id, entry_point
        self.kwargs = dict(kwargs or {})
        self.max_episode_steps = max_episode_steps
        self.reward_threshold = reward_threshold

    def make(self, **kwargs):
        merged = dict(self.kwargs)
        merged.update(kwargs)
        ctor = self.entry_point
        if isinstance(ctor, str):
            mod, attr = ctor.split(":")
            ctor = getattr(importlib.import_module(mod), attr)
        env = ctor(**merged)
        env.unwrapped.spec = self
        # Batched envs count steps per env inside the kernel (A12); only the scalar,
        # reference-shaped env gets the host-side TimeLimit wrapper.
        if self.max_episode_steps is not None and not getattr(env, "is_batched", False):
            env = TimeLimit(env, self.max_episode_steps)
        return env


class _Registry:
    def __init__(self):
        self.env_specs: Dict[str, EnvSpec] = {}

    def register(self, id, **kwargs):
        if id in self.env_specs:
            raise ValueError(f"Cannot re-register id: {id}")
        self.env_specs[id] = EnvSpec(id, **kwargs)

    def spec(self, id) -> EnvSpec:
        try:
            return self.env_specs[id]
        except KeyError:
            raise KeyError(f"No registered env with id: {id}")

    def make(self, id, **kwargs):
        return self.spec(id).make(**kwargs)

    def all(self):
        return self.env_specs.values()


registry = _Registry()


def register(id, **kwargs):
    return registry.register(id, **kwargs)


def make(id, **kwargs):
    return registry.make(id, **kwargs)


def spec(id):
    return registry.spec(id)


# module-shaped namespaces so `gym.spaces.Box`, `gym.envs.register`, `gym.wrappers.TimeLimit` resolve
spaces = types.SimpleNamespace(Box=Box, Space=Space)
envs = types.SimpleNamespace(register=register, registry=registry, make=make, spec=spec)
wrappers = types.SimpleNamespace(TimeLimit=TimeLimit)
core = types.SimpleNamespace(Env=Env, Wrapper=Wrapper, ObsType=Any)
__version__ = "0.20.0-shim"
