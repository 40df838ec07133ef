"""ctypes front-end of tests/host_emu/host_emu.cpp (TEST BUILD ONLY: the engine's per-lane device headers
compiled for the CPU). Same surface as oracle.OracleVecEnv so the parity helpers can drive either."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_LIB = os.path.join(_HERE, "libmg_host_emu.so")
KIND = {"empty": 0, "doorkey": 1, "crossing": 2, "fourrooms": 3, "lavagap": 4, "distshift": 5, "multiroom": 6,
        "lockedroom": 7, "playground": 8,  # 7 and up: device generators checked here before the kernels are instantiated
        "gotodoor": 9, "fetch": 10, "redbluedoors": 11, "gotoobject": 12, "putnear": 13, "memory": 14, "dynobstacles": 15, "roomgrid": 16}
AUTORESET = {"next_step": 0, "same_step": 1, "disabled": 2}
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "host_emu.cpp")
        deps = [src] + [os.path.join(_ROOT, "minigrid_b200", "csrc", f) for f in os.listdir(os.path.join(_ROOT, "minigrid_b200", "csrc"))]
        if not os.path.exists(_LIB) or any(os.path.getmtime(d) > os.path.getmtime(_LIB) for d in deps):
            subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                                   "-ffp-contract=off", "-o", _LIB, src])
        L = C.CDLL(_LIB)
        p = C.c_void_p
        L.emu_create.restype = p
        L.emu_create.argtypes = [C.c_int] * 5 + [p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.emu_destroy.argtypes = [p]
        L.emu_seed.argtypes = [p, p]
        L.emu_reset.argtypes = [p, p, p]
        L.emu_step.restype = C.c_int
        L.emu_step.argtypes = [p] * 7
        L.emu_step_packed.restype = C.c_int
        L.emu_step_packed.argtypes = [p, p, p]
        L.emu_full_obs.argtypes = [p, p, C.c_int]
        L.emu_get_state.argtypes = [p] * 4
        L.emu_set_state.argtypes = [p] * 3
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class EmuVecEnv:
    def __init__(self, spec, num_envs, autoreset="next_step", layout=-1):
        kind, W, H, max_steps, see_through, params = spec
        self.width, self.height, self.num_envs = W, H, int(num_envs)
        self._max_steps = max_steps
        prm = np.asarray(list(params), dtype=np.int32)
        self._h = lib().emu_create(KIND[kind], W, H, max_steps, int(see_through), _ptr(prm), len(prm), self.num_envs,
                                   AUTORESET[autoreset], layout)
        n = self.num_envs
        self.obs = np.zeros((n, 7, 7, 3), np.uint8)
        self.dir = np.zeros(n, np.int32)
        self.reward = np.zeros(n, np.float64)
        self.terminated = np.zeros(n, np.uint8)
        self.truncated = np.zeros(n, np.uint8)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().emu_destroy(self._h)
            self._h = None

    def reset(self, seed=None):
        if seed is not None:
            seeds = np.arange(self.num_envs, dtype=np.uint64) + np.uint64(seed) if np.isscalar(seed) else np.asarray(seed, np.uint64)
            lib().emu_seed(self._h, _ptr(np.ascontiguousarray(seeds)))
        lib().emu_reset(self._h, _ptr(self.obs), _ptr(self.dir))
        return self.obs, self.dir

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.int32)
        rc = lib().emu_step(self._h, _ptr(a), _ptr(self.obs), _ptr(self.dir), _ptr(self.reward), _ptr(self.terminated), _ptr(self.truncated))
        if rc != 0:
            raise ValueError("Unknown action")
        return self.obs, self.dir, self.reward, self.terminated.astype(bool), self.truncated.astype(bool)

    def step_packed(self, actions):
        """The packed host path: K1's 52-byte records (device header code), expanded by the PRODUCT's host expander
        (mg_expand_packed in libminigrid_b200.so: host code, no GPU needed)."""
        from minigrid_b200 import _lib

        a = np.ascontiguousarray(actions, dtype=np.int32)
        packed = np.zeros((self.num_envs, 52), np.uint8)
        rc = lib().emu_step_packed(self._h, _ptr(a), _ptr(packed))
        if rc != 0:
            raise ValueError("Unknown action")
        L = _lib.load()
        max_steps = self._max_steps
        rc = L.mg_expand_packed(_ptr(packed), self.num_envs, max_steps, _ptr(self.obs), _ptr(self.dir), _ptr(self.reward),
                                _ptr(self.terminated), _ptr(self.truncated))
        assert rc == 0
        return self.obs, self.dir, self.reward, self.terminated.astype(bool), self.truncated.astype(bool)

    def gen_obs(self):
        # an obs-only pass is a step with actions == NULL
        lib().emu_step(self._h, None, _ptr(self.obs), _ptr(self.dir), None, None, None)
        return self.obs, self.dir

    def full_obs(self):
        out = np.zeros((self.num_envs, self.width, self.height, 3), np.uint8)
        lib().emu_full_obs(self._h, _ptr(out), 1)
        return out

    def get_state(self):
        n = self.num_envs
        st = {"grid": np.zeros((n, self.width, self.height, 3), np.uint8), "agent": np.zeros((n, 6), np.int32),
              "rng": np.zeros((n, 6), np.uint64), "pending": np.zeros(n, np.uint8)}
        lib().emu_full_obs(self._h, _ptr(st["grid"]), 0)
        lib().emu_get_state(self._h, _ptr(st["agent"]), _ptr(st["rng"]), _ptr(st["pending"]))
        return st

    def set_state(self, grid=None, agent=None):
        g = None if grid is None else np.ascontiguousarray(grid, np.uint8)
        a = None if agent is None else np.ascontiguousarray(agent, np.int32)
        lib().emu_set_state(self._h, _ptr(g), _ptr(a))
