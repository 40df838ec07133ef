"""Batched equivalents of the two reference wrappers BASELINE.json's north_star names.

ImgObsWrapper   (minigrid/wrappers.py:187-214): the observation is obs["image"].
FullyObsWrapper (minigrid/wrappers.py:383-426): obs["image"] = grid.encode() with the agent cell set to
                (OBJECT_TO_IDX["agent"], COLOR_TO_IDX["red"], agent_dir); shape (W, H, 3) per env.
"""
from __future__ import annotations

import torch


class _VecWrapper:
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith("_") or name == "env":
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return getattr(self.env, "unwrapped", self.env)

    def observation(self, obs):
        raise NotImplementedError

    def reset(self, *, seed=None, options=None):
        obs, info = self.env.reset(seed=seed, options=options)
        return self.observation(obs), info

    def step(self, actions):
        obs, r, te, tr, info = self.env.step(actions)
        return self.observation(obs), r, te, tr, info

    def close(self):
        return self.env.close()


class ImgObsWrapper(_VecWrapper):
    def observation(self, obs):
        return obs["image"]


class FullyObsWrapper(_VecWrapper):
    def __init__(self, env):
        super().__init__(env)
        base = self.unwrapped
        self._full = torch.empty((base.num_envs, base.width, base.height, 3), dtype=torch.uint8, device=base.device)

    def observation(self, obs):
        self.unwrapped.full_obs(out=self._full)
        return {**obs, "image": self._full}
