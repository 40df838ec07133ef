"""Env / Wrapper stand-ins following the public gymnasium API the reference subclasses."""
from __future__ import annotations

from typing import Any, TypeVar

from .utils import seeding

ObsType = TypeVar("ObsType")
ActType = TypeVar("ActType")
WrapperObsType = TypeVar("WrapperObsType")
WrapperActType = TypeVar("WrapperActType")


class Env:
    metadata: dict = {"render_modes": []}
    render_mode = None
    spec = None
    _np_random = None
    _np_random_seed = None

    def __class_getitem__(cls, item):
        return cls

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self._np_random, self._np_random_seed = seeding.np_random(seed)

    def step(self, action):
        raise NotImplementedError

    def render(self):
        raise NotImplementedError

    def close(self):
        pass

    @property
    def unwrapped(self):
        return self

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random, self._np_random_seed = seeding.np_random()
        return self._np_random

    @np_random.setter
    def np_random(self, value):
        self._np_random = value

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
        return False


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self._action_space = None
        self._observation_space = None

    def __class_getitem__(cls, item):
        return cls

    def __getattr__(self, name):
        if name.startswith("_") or name == "env":
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def action_space(self):
        return self._action_space if self._action_space is not None else self.env.action_space

    @action_space.setter
    def action_space(self, v):
        self._action_space = v

    @property
    def observation_space(self):
        return self._observation_space if self._observation_space is not None else self.env.observation_space

    @observation_space.setter
    def observation_space(self, v):
        self._observation_space = v

    @property
    def unwrapped(self):
        return self.env.unwrapped

    @property
    def np_random(self):
        return self.env.np_random

    @property
    def render_mode(self):
        return self.env.render_mode

    def reset(self, *, seed=None, options=None):
        return self.env.reset(seed=seed, options=options)

    def step(self, action):
        return self.env.step(action)

    def render(self):
        return self.env.render()

    def close(self):
        return self.env.close()


class ObservationWrapper(Wrapper):
    def reset(self, *, seed=None, options=None):
        obs, info = self.env.reset(seed=seed, options=options)
        return self.observation(obs), info

    def step(self, action):
        obs, r, term, trunc, info = self.env.step(action)
        return self.observation(obs), r, term, trunc, info

    def observation(self, obs):
        raise NotImplementedError


class ActionWrapper(Wrapper):
    def step(self, action):
        return self.env.step(self.action(action))

    def action(self, action):
        raise NotImplementedError


class RewardWrapper(Wrapper):
    def step(self, action):
        obs, r, term, trunc, info = self.env.step(action)
        return obs, self.reward(r), term, trunc, info

    def reward(self, r):
        raise NotImplementedError
