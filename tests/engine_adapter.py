"""Drives the CUDA engine (through the Python host layer -> C-ABI) with the same surface as
oracle.OracleVecEnv, returning numpy arrays, so tests/parity.py can compare the two."""
import numpy as np
import torch

from minigrid_b200 import MinigridVecEnv, specs


class EngineAdapter:
    def __init__(self, env_id=None, num_envs=1, mode="next_step", spec=None, host=False, action_dtype=torch.int32,
                 host_format="full", host_threads=0):
        self.e = MinigridVecEnv(env_id, num_envs, spec=spec, autoreset_mode=mode)
        if host_format != "full":
            self.e.set_host_format(host_format, host_threads)
        self.num_envs = num_envs
        self.host = host
        self.action_dtype = action_dtype

    @staticmethod
    def _np(t):
        return t.cpu().numpy() if isinstance(t, torch.Tensor) else t

    def reset(self, seed=None, mask=None):
        if self.host:
            obs, _ = self.e.reset_host(seed=seed)
        else:
            obs, _ = self.e.reset(seed=seed, options=None if mask is None else {"reset_mask": mask})
        return self._np(obs["image"]).copy(), self._np(obs["direction"]).copy()

    def step(self, actions):
        if self.host:
            obs, r, te, tr, _ = self.e.step_host(np.asarray(actions, np.int32))
        else:
            a = actions if isinstance(actions, torch.Tensor) else torch.as_tensor(np.asarray(actions))
            a = a.to(device=self.e.device, dtype=self.action_dtype)
            obs, r, te, tr, _ = self.e.step(a)
            self.e.check_actions()
        return (self._np(obs["image"]).copy(), self._np(obs["direction"]).copy(), self._np(r).copy(),
                self._np(te).copy(), self._np(tr).copy())

    def set_no_death(self, no_death_types, death_cost=-1.0):
        self.e.set_no_death(tuple(no_death_types), death_cost)

    def set_bonus(self, kind):
        self.e.set_bonus(kind)

    def gen_obs(self):
        obs = self.e.gen_obs()
        return self._np(obs["image"]).copy(), self._np(obs["direction"]).copy()

    def full_obs(self):
        return self.e.full_obs().cpu().numpy()

    def get_state(self):
        st = {k: v.cpu().numpy() for k, v in self.e.get_state().items()}
        st["rng"] = st["rng"].view(np.uint64)
        return st

    def set_state(self, grid=None, agent=None, rng=None, pending=None):
        self.e.set_state(grid=grid, agent=agent, rng=rng, pending=pending)


def make_engine(env_id, n, mode):
    return EngineAdapter(env_id, n, mode)


def spec_tuple(env_id):
    s = specs.get(env_id)
    kind = ["empty", "doorkey", "crossing", "fourrooms", "lavagap", "distshift", "multiroom", "lockedroom", "playground",
            "gotodoor", "fetch", "redbluedoors", "gotoobject", "putnear", "memory", "dynobstacles", "roomgrid"][s.kind]
    return (kind, s.width, s.height, s.max_steps, s.see_through_walls, list(s.params))
