"""ctypes front-end of the C oracle (oracle/mg_oracle.c). TEST INFRASTRUCTURE ONLY.

`OracleVecEnv` mirrors a list of reference `MiniGridEnv` objects driven in lockstep with
gymnasium.vector.SyncVectorEnv autoreset semantics (restated in mg_oracle.c:step_range).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmg_oracle.so")

KIND = {"empty": 0, "doorkey": 1, "crossing": 2, "fourrooms": 3, "lavagap": 4, "distshift": 5, "multiroom": 6,
        "lockedroom": 7, "playground": 8, "gotodoor": 9, "fetch": 10, "redbluedoors": 11, "gotoobject": 12, "putnear": 13,
        "memory": 14, "dynobstacles": 15, "roomgrid": 16}
AUTORESET = {"next_step": 0, "same_step": 1, "disabled": 2}

def _om(variant, rows, cols, visited, key_in_box, blocked, agent_room=(0, 0), quarters=0):
    """envs/obstructedmaze.py:79-105: room_size 6, max_steps = 4 * num_rooms_visited * 36"""
    return ("roomgrid", 5 * cols + 1, 5 * rows + 1, 4 * visited * 36, False,
            [variant, 6, rows, cols, int(key_in_box), int(blocked), agent_room[0] | (agent_room[1] << 4), quarters])


# id -> (kind, width, height, max_steps, see_through_walls, params); restated from
# /root/reference/minigrid/__init__.py:25-28,106-109,159-162,183-185,214-216 and the env
# constructors (empty.py:69-90, doorkey.py:62-68, crossing.py:89-116, fourrooms.py:55-67).
ENV_SPECS = {
    "MiniGrid-Empty-5x5-v0": ("empty", 5, 5, 100, True, [0, 1, 1, 0]),
    "MiniGrid-Empty-8x8-v0": ("empty", 8, 8, 256, True, [0, 1, 1, 0]),
    "MiniGrid-DoorKey-8x8-v0": ("doorkey", 8, 8, 640, False, []),
    "MiniGrid-LavaCrossingS9N1-v0": ("crossing", 9, 9, 324, False, [1, 9]),
    "MiniGrid-FourRooms-v0": ("fourrooms", 19, 19, 100, False, []),
    # extra ids of the same four generators (wider parity coverage)
    "MiniGrid-Empty-Random-6x6-v0": ("empty", 6, 6, 144, True, [1, 0, 0, 0]),
    "MiniGrid-Empty-16x16-v0": ("empty", 16, 16, 1024, True, [0, 1, 1, 0]),
    "MiniGrid-DoorKey-5x5-v0": ("doorkey", 5, 5, 250, False, []),
    "MiniGrid-DoorKey-16x16-v0": ("doorkey", 16, 16, 2560, False, []),
    "MiniGrid-LavaCrossingS9N3-v0": ("crossing", 9, 9, 324, False, [3, 9]),
    "MiniGrid-LavaCrossingS11N5-v0": ("crossing", 11, 11, 484, False, [5, 9]),
    "MiniGrid-SimpleCrossingS9N2-v0": ("crossing", 9, 9, 324, False, [2, 2]),
    # round-1 widening: lavagap.py:68-135 (__init__.py:295-310), distshift.py:63-120 (__init__.py:79-88)
    "MiniGrid-LavaGapS5-v0": ("lavagap", 5, 5, 100, False, [9]),
    "MiniGrid-LavaGapS7-v0": ("lavagap", 7, 7, 196, False, [9]),
    "MiniGrid-DistShift1-v0": ("distshift", 9, 7, 252, True, [2, 1, 1, 0]),
    "MiniGrid-DistShift2-v0": ("distshift", 9, 7, 252, True, [5, 1, 1, 0]),
    # multiroom.py:77-115 (25x25, max_steps = maxNumRooms * 20), __init__.py:363-385
    "MiniGrid-MultiRoom-N2-S4-v0": ("multiroom", 25, 25, 40, False, [2, 2, 4]),
    "MiniGrid-MultiRoom-N4-S5-v0": ("multiroom", 25, 25, 120, False, [6, 6, 5]),
    "MiniGrid-MultiRoom-N6-v0": ("multiroom", 25, 25, 120, False, [6, 6, 10]),
    "MiniGrid-LockedRoom-v0": ("lockedroom", 19, 19, 190, False, []),
    "MiniGrid-Playground-v0": ("playground", 19, 19, 100, False, []),
    "MiniGrid-GoToDoor-5x5-v0": ("gotodoor", 5, 5, 100, True, []),
    "MiniGrid-GoToDoor-6x6-v0": ("gotodoor", 6, 6, 144, True, []),
    "MiniGrid-GoToDoor-8x8-v0": ("gotodoor", 8, 8, 256, True, []),
    "MiniGrid-Fetch-5x5-N2-v0": ("fetch", 5, 5, 125, True, [2]),
    "MiniGrid-Fetch-6x6-N2-v0": ("fetch", 6, 6, 180, True, [2]),
    "MiniGrid-Fetch-8x8-N3-v0": ("fetch", 8, 8, 320, True, [3]),
    "MiniGrid-RedBlueDoors-6x6-v0": ("redbluedoors", 12, 6, 720, False, []),
    "MiniGrid-RedBlueDoors-8x8-v0": ("redbluedoors", 16, 8, 1280, False, []),
    "MiniGrid-GoToObject-6x6-N2-v0": ("gotoobject", 6, 6, 180, True, [2]),
    "MiniGrid-GoToObject-8x8-N2-v0": ("gotoobject", 8, 8, 320, True, [2]),
    "MiniGrid-PutNear-6x6-N2-v0": ("putnear", 6, 6, 30, True, [2]),
    "MiniGrid-PutNear-8x8-N3-v0": ("putnear", 8, 8, 40, True, [3]),
    "MiniGrid-MemoryS17Random-v0": ("memory", 17, 17, 1445, False, [1]),
    "MiniGrid-MemoryS13Random-v0": ("memory", 13, 13, 845, False, [1]),
    "MiniGrid-MemoryS13-v0": ("memory", 13, 13, 845, False, [0]),
    "MiniGrid-MemoryS11-v0": ("memory", 11, 11, 605, False, [0]),
    "MiniGrid-MemoryS9-v0": ("memory", 9, 9, 405, False, [0]),
    "MiniGrid-MemoryS7-v0": ("memory", 7, 7, 245, False, [0]),
    "MiniGrid-Dynamic-Obstacles-5x5-v0": ("dynobstacles", 5, 5, 100, True, [2, 0, 1, 1, 0]),
    "MiniGrid-Dynamic-Obstacles-Random-5x5-v0": ("dynobstacles", 5, 5, 100, True, [2, 1, 0, 0, 0]),
    "MiniGrid-Dynamic-Obstacles-6x6-v0": ("dynobstacles", 6, 6, 144, True, [3, 0, 1, 1, 0]),
    "MiniGrid-Dynamic-Obstacles-Random-6x6-v0": ("dynobstacles", 6, 6, 144, True, [3, 1, 0, 0, 0]),
    "MiniGrid-Dynamic-Obstacles-8x8-v0": ("dynobstacles", 8, 8, 256, True, [4, 0, 1, 1, 0]),
    "MiniGrid-Dynamic-Obstacles-16x16-v0": ("dynobstacles", 16, 16, 1024, True, [8, 0, 1, 1, 0]),
    "MiniGrid-Unlock-v0": ("roomgrid", 11, 6, 288, False, [0, 6, 1, 2]),
    "MiniGrid-UnlockPickup-v0": ("roomgrid", 11, 6, 288, False, [1, 6, 1, 2]),
    "MiniGrid-BlockedUnlockPickup-v0": ("roomgrid", 11, 6, 576, False, [2, 6, 1, 2]),
    "MiniGrid-KeyCorridorS3R1-v0": ("roomgrid", 7, 3, 270, False, [3, 3, 1, 3]),
    "MiniGrid-KeyCorridorS3R2-v0": ("roomgrid", 7, 5, 270, False, [3, 3, 2, 3]),
    "MiniGrid-KeyCorridorS3R3-v0": ("roomgrid", 7, 7, 270, False, [3, 3, 3, 3]),
    "MiniGrid-KeyCorridorS4R3-v0": ("roomgrid", 10, 10, 480, False, [3, 4, 3, 3]),
    "MiniGrid-KeyCorridorS5R3-v0": ("roomgrid", 13, 13, 750, False, [3, 5, 3, 3]),
    "MiniGrid-KeyCorridorS6R3-v0": ("roomgrid", 16, 16, 1080, False, [3, 6, 3, 3]),
    "MiniGrid-ObstructedMaze-1Dl-v0": _om(4, 1, 2, 2, False, False),
    "MiniGrid-ObstructedMaze-1Dlh-v0": _om(4, 1, 2, 2, True, False),
    "MiniGrid-ObstructedMaze-1Dlhb-v0": _om(4, 1, 2, 2, True, True),
    "MiniGrid-ObstructedMaze-2Dl-v0": _om(5, 3, 3, 4, False, False, (2, 1), 1),
    "MiniGrid-ObstructedMaze-2Dlh-v0": _om(5, 3, 3, 4, True, False, (2, 1), 1),
    "MiniGrid-ObstructedMaze-2Dlhb-v0": _om(5, 3, 3, 4, True, True, (2, 1), 1),
    "MiniGrid-ObstructedMaze-1Q-v0": _om(5, 3, 3, 5, True, True, (1, 1), 1),
    "MiniGrid-ObstructedMaze-2Q-v0": _om(5, 3, 3, 11, True, True, (2, 1), 2),
    "MiniGrid-ObstructedMaze-Full-v0": _om(5, 3, 3, 25, True, True, (1, 1), 4),
    "MiniGrid-ObstructedMaze-2Dlhb-v1": _om(6, 3, 3, 4, True, True, (2, 1), 1),
    "MiniGrid-ObstructedMaze-1Q-v1": _om(6, 3, 3, 5, True, True, (1, 1), 1),
    "MiniGrid-ObstructedMaze-2Q-v1": _om(6, 3, 3, 11, True, True, (2, 1), 2),
    "MiniGrid-ObstructedMaze-Full-v1": _om(6, 3, 3, 25, True, True, (1, 1), 4),
}

# SURVEY 8(f-1) generators restated ahead of their device kernels: the oracle and its fixtures exist, the product does
# not register these ids yet (lockedroom.py:74-90 / __init__.py:312-318, playground.py:16-25 / __init__.py:516-522)
NEXT_SPECS = {
    # envs/obstructedmaze.py + obstructedmaze_v1.py, __init__.py:387-514
    # SURVEY 8(f-2), second half: core/roomgrid.py + unlock.py:55-70 (2 rooms of 6, 8 * 36 steps), unlockpickup.py:60-78,
    # blockedunlockpickup.py:67-85 (16 * 36), keycorridor.py:73-97 (3 columns, 30 * room_size^2); __init__.py:12-20, 252-290, 555-563
    # params {variant, room_size, num_rows, num_cols}
    # SURVEY 8(f-2), first of the step post-filters: gotodoor.py:65-86 (4 * size^2 steps, see_through_walls=True), __init__.py:218-236
    # fetch.py:72-103 (5 * size^2 steps, see_through_walls=True), __init__.py:196-208; params {numObjs}
    # redbluedoors.py:60-72 (2 size x size, 20 * size^2 steps), __init__.py:541-551
    # gotoobject.py:66-90 (size 6, numObjs 2, 5 * size^2 steps, see_through_walls=True), __init__.py:241-250; params {numObjs}
    # putnear.py:66-92 (size 6, numObjs 2, 5 * size steps, see_through_walls=True), __init__.py:527-537; params {numObjs}
    # memory.py:67-88 (5 * size^2 steps, see_through_walls=False), __init__.py:323-357; params {random_length}
    # dynamicobstacles.py:72-105 (4 * size^2 steps, see_through_walls=True), __init__.py:117-153;
    # params {n_obstacles, random_start, start_x, start_y, start_dir}
}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mg_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmg_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        p = C.c_void_p
        L.mgo_vec_create.restype = p
        L.mgo_vec_create.argtypes = [C.c_int] * 5 + [p, C.c_int, C.c_int]
        L.mgo_vec_destroy.argtypes = [p]
        L.mgo_vec_set_no_death.argtypes = [p, C.c_int, C.c_double]
        L.mgo_vec_set_bonus.argtypes = [p, C.c_int]
        L.mgo_vec_seed.argtypes = [p, p]
        L.mgo_vec_reset.argtypes = [p, p, p, C.c_int]
        L.mgo_vec_seed_masked.argtypes = [p, p, p]
        L.mgo_vec_reset_masked.argtypes = [p, p, p, p]
        L.mgo_vec_step.restype = C.c_int
        L.mgo_vec_step.argtypes = [p] * 7 + [C.c_int, C.c_int]
        L.mgo_vec_full_obs.argtypes = [p, p]
        L.mgo_vec_gen_obs.argtypes = [p, p, p]
        L.mgo_vec_gen_obs_view.restype = C.c_int
        L.mgo_vec_gen_obs_view.argtypes = [p, C.c_int, p]
        L.mgo_vec_symbolic_obs.argtypes = [p, p]
        L.mgo_vec_get_state.argtypes = [p] * 5
        L.mgo_vec_set_state.argtypes = [p] * 5
        L.mgo_rng_integers.restype = C.c_int64
        L.mgo_rng_integers.argtypes = [p, C.c_int, C.c_int64, C.c_int64]
        L.mgo_rng_next32.restype = C.c_uint32
        L.mgo_rng_next32.argtypes = [p, C.c_int]
        L.mgo_rng_shuffle_perm.argtypes = [p, C.c_int, p, C.c_int]
        L.mgo_vec_rollout.restype = C.c_double
        L.mgo_vec_rollout.argtypes = [p, p, C.c_int, C.c_int, C.c_int, p]
        L.mgo_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleVecEnv:
    def __init__(self, env_id=None, num_envs=1, *, spec=None, autoreset="next_step", n_threads=1):
        kind, W, H, max_steps, see_through, params = spec if spec is not None else {**ENV_SPECS, **NEXT_SPECS}[env_id]
        self.kind, self.width, self.height = kind, W, H
        self.max_steps, self.see_through, self.params = max_steps, see_through, list(params)
        self.num_envs = int(num_envs)
        self.mode = AUTORESET[autoreset]
        self.n_threads = n_threads
        prm = np.asarray(self.params, dtype=np.int32)
        self._h = lib().mgo_vec_create(KIND[kind], W, H, max_steps, int(see_through), _ptr(prm), len(prm), self.num_envs)
        n = self.num_envs
        self.obs = np.zeros((n, 7, 7, 3), np.uint8)
        self.dir = np.zeros(n, np.int32)
        self.reward = np.zeros(n, np.float64)
        self.terminated = np.zeros(n, np.uint8)
        self.truncated = np.zeros(n, np.uint8)

    def close(self):
        if self._h:
            lib().mgo_vec_destroy(self._h)
            self._h = None

    __del__ = close

    def seed(self, seeds):
        s = np.ascontiguousarray(seeds, dtype=np.uint64)
        assert s.shape == (self.num_envs,)
        lib().mgo_vec_seed(self._h, _ptr(s))

    def reset(self, seed=None, mask=None):
        m = None if mask is None else np.ascontiguousarray(np.asarray(mask) != 0, dtype=np.uint8)
        if seed is not None:
            seeds = np.arange(self.num_envs, dtype=np.uint64) + np.uint64(seed) if np.isscalar(seed) else seed
            if m is None:
                self.seed(seeds)
            else:
                s = np.ascontiguousarray(seeds, dtype=np.uint64)
                lib().mgo_vec_seed_masked(self._h, _ptr(m), _ptr(s))
        if m is None:
            lib().mgo_vec_reset(self._h, _ptr(self.obs), _ptr(self.dir), self.n_threads)
        else:
            lib().mgo_vec_reset_masked(self._h, _ptr(m), _ptr(self.obs), _ptr(self.dir))
        return self.obs, self.dir

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.int32)
        assert a.shape == (self.num_envs,)
        rc = lib().mgo_vec_step(self._h, _ptr(a), _ptr(self.obs), _ptr(self.dir), _ptr(self.reward),
                                _ptr(self.terminated), _ptr(self.truncated), self.mode, self.n_threads)
        if rc != 0:
            raise ValueError("Unknown action")
        return self.obs, self.dir, self.reward, self.terminated.astype(bool), self.truncated.astype(bool)

    # ---- reward wrappers around every env (wrappers.py:68-184, 809-882); the bonus wrapper is the outermost ----
    OBJECT_TO_IDX = {"unseen": 0, "empty": 1, "wall": 2, "floor": 3, "door": 4, "key": 5, "ball": 6, "box": 7, "goal": 8, "lava": 9, "agent": 10}

    def set_no_death(self, no_death_types, death_cost=-1.0):
        """NoDeath(env, no_death_types, death_cost); an empty tuple removes the wrapper."""
        mask = 0
        for t in no_death_types:
            mask |= 1 << self.OBJECT_TO_IDX[t]
        if lib().mgo_vec_set_no_death(self._h, mask, float(death_cost)) != 0:
            raise AssertionError("goal cannot be a death cell")

    def set_bonus(self, kind):
        """kind: None, "action" (ActionBonus) or "position" (PositionBonus): a fresh wrapper with empty counts."""
        if lib().mgo_vec_set_bonus(self._h, {None: 0, "action": 1, "position": 2}[kind]) != 0:
            raise MemoryError("bonus counts")

    def gen_obs(self):
        lib().mgo_vec_gen_obs(self._h, _ptr(self.obs), _ptr(self.dir))
        return self.obs, self.dir

    def gen_obs_view(self, view_size):
        """ViewSizeWrapper.observation: uint8[n, V, V, 3]."""
        out = np.zeros((self.num_envs, view_size, view_size, 3), np.uint8)
        if lib().mgo_vec_gen_obs_view(self._h, int(view_size), _ptr(out)) != 0:
            raise ValueError("view size must be odd and in 3..15")
        return out

    def symbolic_obs(self):
        """SymbolicObsWrapper.observation: int64[n, W, H, 3]."""
        out = np.zeros((self.num_envs, self.width, self.height, 3), np.int64)
        lib().mgo_vec_symbolic_obs(self._h, _ptr(out))
        return out

    @staticmethod
    def one_hot(image):
        """OneHotPartialObsWrapper.observation (wrappers.py:268-284) on uint8[..., 3] images: 11 + 6 + 3 channels."""
        img = np.asarray(image)
        out = np.zeros(img.shape[:-1] + (20,), np.uint8)
        idx = np.indices(img.shape[:-1])
        out[(*idx, img[..., 0])] = 1
        out[(*idx, 11 + img[..., 1])] = 1
        out[(*idx, 17 + img[..., 2])] = 1
        return out

    def full_obs(self):
        out = np.zeros((self.num_envs, self.width, self.height, 3), np.uint8)
        lib().mgo_vec_full_obs(self._h, _ptr(out))
        return out

    def get_state(self):
        n = self.num_envs
        st = {
            "grid": np.zeros((n, self.width, self.height, 3), np.uint8),
            "agent": np.zeros((n, 6), np.int32),
            "rng": np.zeros((n, 6), np.uint64),
            "pending": np.zeros(n, np.uint8),
        }
        lib().mgo_vec_get_state(self._h, _ptr(st["grid"]), _ptr(st["agent"]), _ptr(st["rng"]), _ptr(st["pending"]))
        return st

    def set_state(self, grid=None, agent=None, rng=None, pending=None):
        g = None if grid is None else np.ascontiguousarray(grid, np.uint8)
        a = None if agent is None else np.ascontiguousarray(agent, np.int32)
        r = None if rng is None else np.ascontiguousarray(rng, np.uint64)
        p = None if pending is None else np.ascontiguousarray(pending, np.uint8)
        lib().mgo_vec_set_state(self._h, _ptr(g), _ptr(a), _ptr(r), _ptr(p))

    def rng_integers(self, i, low, high):
        return int(lib().mgo_rng_integers(self._h, i, low, high))

    def rng_next32(self, i):
        return int(lib().mgo_rng_next32(self._h, i))

    def rng_shuffle(self, i, n):
        perm = np.arange(n, dtype=np.int32)
        lib().mgo_rng_shuffle_perm(self._h, i, _ptr(perm), n)
        return perm

    def rollout(self, actions, n_threads=0):
        """actions int32[T, N]; returns (seconds, checksum of the last obs batch)."""
        a = np.ascontiguousarray(actions, dtype=np.int32)
        assert a.ndim == 2 and a.shape[1] == self.num_envs
        ck = C.c_uint64(0)
        secs = lib().mgo_vec_rollout(self._h, _ptr(a), a.shape[0], self.mode, n_threads, C.byref(ck))
        return float(secs), int(ck.value)


def max_threads():
    return int(lib().mgo_max_threads())
