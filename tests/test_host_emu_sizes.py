"""Grid sizes the registered ids never use (3-wide corridors, 26 x 26 = the engine's limit, strongly non-square), with
random object soups, through the device headers compiled for the CPU (tests/host_emu) in BOTH HBM layouts, against
the oracle. The oracle's transition / gen_obs do not depend on the size, and are pinned to the reference at the
registered sizes (tests/test_oracle_golden.py)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emu"))
from emu import EmuVecEnv  # noqa: E402

from oracle.oracle import OracleVecEnv  # noqa: E402


def random_soup(rng, W, H):
    """encoded grid [W][H][3]: grey border, interior of every object kind, colour and door state"""
    g = np.zeros((W, H, 3), np.uint8)
    g[:, :, 0] = 1
    for x in range(W):
        for y in range(H):
            if x in (0, W - 1) or y in (0, H - 1):
                g[x, y] = (2, 5, 0)
            elif rng.random() < 0.45:
                t = int(rng.choice([2, 3, 4, 4, 4, 5, 6, 7, 8, 9]))
                col = 1 if t == 8 else 0 if t == 9 else int(rng.integers(0, 6))
                g[x, y] = (t, col, int(rng.integers(0, 3)) if t == 4 else 0)
    return g


SIZES = [(3, 3), (3, 26), (26, 3), (4, 5), (26, 26), (25, 26), (26, 7), (9, 21), (16, 16), (5, 26)]


@pytest.mark.parametrize("layout", [0, 1], ids=["tiled", "window"])
@pytest.mark.parametrize("W,H", SIZES, ids=[f"{w}x{h}" for w, h in SIZES])
def test_soups_at_unusual_sizes(W, H, layout):
    rng = np.random.default_rng(W * 100 + H)
    n, steps = 37, 90  # 37: one full tile and a ragged one
    see_through = bool((W + H) & 1)
    spec = ("empty", W, H, 60, see_through, [0, 1, 1, 0])
    emu = EmuVecEnv(spec, n, autoreset="disabled", layout=layout)
    orc = OracleVecEnv(None, n, autoreset="disabled", spec=spec)
    emu.reset(seed=0)
    orc.reset(seed=0)
    grid = np.stack([random_soup(rng, W, H) for _ in range(n)])
    agent = np.zeros((n, 6), np.int32)
    for i in range(n):
        while True:  # a cell the agent could legally stand on
            ax, ay = int(rng.integers(1, W - 1)), int(rng.integers(1, H - 1))
            t, _, s = grid[i, ax, ay]
            if t in (1, 3, 8, 9) or (t == 4 and s == 0):
                break
            grid[i, ax, ay] = (1, 0, 0)
        carry_t = int(rng.choice([-1, 5, 6, 7]))
        agent[i] = [ax, ay, int(rng.integers(0, 4)), carry_t, int(rng.integers(0, 6)) if carry_t >= 0 else 0, int(rng.integers(0, 40))]
    emu.set_state(grid=grid, agent=agent)
    orc.set_state(grid=grid, agent=agent)
    a, b = emu.gen_obs(), orc.gen_obs()
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])
    np.testing.assert_array_equal(emu.full_obs(), orc.full_obs())
    for t in range(steps):
        act = rng.integers(0, 7, n).astype(np.int32)
        x, y = emu.step(act), orc.step(act)
        for u, v, name in zip(x, y, ["obs", "dir", "reward", "terminated", "truncated"]):
            np.testing.assert_array_equal(np.asarray(u), np.asarray(v), err_msg=f"{name} t={t}")
    s1, s2 = emu.get_state(), orc.get_state()
    np.testing.assert_array_equal(s1["grid"], s2["grid"])
    np.testing.assert_array_equal(s1["agent"], s2["agent"])
