"""the observation wrappers against fixtures written by the reference's own wrapper classes
(oracle/gen_golden.py wrappers -> tests/golden/wrappers_*.npz). CPU: the oracle's restatements and the tile-atlas
assembly (the arithmetic of mg_wrappers.cu's RGB kernels, restated in numpy); GPU: the engine's device wrappers."""
import os

import numpy as np
import pytest
from conftest import golden_files, load_golden

from oracle.oracle import OracleVecEnv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURES = golden_files("wrappers")


def _atlas():
    d = np.load(os.path.join(ROOT, "minigrid_b200", "data", "tile_atlas.npz"))
    return d["tiles"], d["index"]


def _code(t, c, s):
    """encode_cell of mg_common.cuh"""
    if t in (0, 1) or t >= 10:
        return 1
    if t == 4:
        return [4, 11, 12][s] | (c << 4) | ((s != 0) << 7)
    return t | (c << 4) | ((t == 2) << 7)


def _oracle_at(g):
    n = g["agent"].shape[0]
    orc = OracleVecEnv(g["env_id"], n)
    orc.reset(seed=0)
    orc.set_state(grid=g["grid"], agent=g["agent"])
    return orc


@pytest.mark.parametrize("path", FIXTURES, ids=os.path.basename)
def test_oracle_wrappers_match_reference_fixture(path):
    g = load_golden(path)
    orc = _oracle_at(g)
    np.testing.assert_array_equal(orc.gen_obs()[0], g["obs"])
    for V in (3, 5, 9, 11):
        np.testing.assert_array_equal(orc.gen_obs_view(V), g[f"view{V}"], err_msg=f"view {V}")
    np.testing.assert_array_equal(orc.symbolic_obs(), g["symbolic"])
    np.testing.assert_array_equal(OracleVecEnv.one_hot(g["obs"]), g["one_hot"])


@pytest.mark.parametrize("path", FIXTURES, ids=os.path.basename)
def test_tile_atlas_reassembles_the_reference_frames(path):
    """What k_rgb_partial / k_rgb_full do, in numpy: tile rows copied out of the atlas the reference rendered."""
    g = load_golden(path)
    tiles, index = _atlas()
    n = g["agent"].shape[0]
    W, H = g["grid"].shape[1:3]
    part = np.zeros((n, 56, 56, 3), np.uint8)
    full = np.zeros((n, H * 8, W * 8, 3), np.uint8)
    for e in range(n):
        ax, ay, d = g["agent"][e, :3]
        dx, dy = [(1, 0), (0, 1), (-1, 0), (0, -1)][d]
        for vx in range(7):
            for vy in range(7):
                t, c, s = (int(v) for v in g["obs"][e, vx, vy])
                seen = t != 0
                code = _code(t, c, s) if seen else 1
                agent = 4 if (vx, vy) == (3, 6) else 0
                part[e, vy * 8:(vy + 1) * 8, vx * 8:(vx + 1) * 8] = tiles[index[code & 0x7F, agent, int(seen)]]
        for x in range(W):
            for y in range(H):
                ox, oy = x - ax, y - ay
                f, l = ox * dx + oy * dy, ox * (-dy) + oy * dx
                vx, vy = l + 3, 6 - f
                hl = 0 <= vx < 7 and 0 <= vy < 7 and g["obs"][e, vx, vy, 0] != 0
                t, c, s = (int(v) for v in g["grid"][e, x, y])
                agent = 1 + d if (x, y) == (ax, ay) else 0
                full[e, y * 8:(y + 1) * 8, x * 8:(x + 1) * 8] = tiles[index[_code(t, c, s) & 0x7F, agent, int(hl)]]
    np.testing.assert_array_equal(part, g["rgb_partial"])
    np.testing.assert_array_equal(full, g["rgb_full"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=os.path.basename)
def test_device_wrappers_match_reference_fixture(path):
    import torch

    import minigrid_b200 as mb

    g = load_golden(path)
    n = g["agent"].shape[0]
    env = mb.MinigridVecEnv(g["env_id"], n, autoreset_mode="disabled")
    env.reset(seed=0)
    env.set_state(grid=g["grid"], agent=g["agent"])
    obs = env.gen_obs()
    np.testing.assert_array_equal(obs["image"].cpu().numpy(), g["obs"])
    for V in (3, 5, 9, 11):
        w = mb.ViewSizeWrapper(env, agent_view_size=V)
        np.testing.assert_array_equal(w.observation(obs)["image"].cpu().numpy(), g[f"view{V}"], err_msg=f"view {V}")
    np.testing.assert_array_equal(mb.ViewSizeWrapper(env, 7).observation(obs)["image"].cpu().numpy(), g["obs"])
    np.testing.assert_array_equal(mb.OneHotPartialObsWrapper(env).observation(obs)["image"].cpu().numpy(), g["one_hot"])
    v5 = mb.ViewSizeWrapper(env, 5).observation(obs)
    np.testing.assert_array_equal(mb.OneHotPartialObsWrapper(env).observation(v5)["image"].cpu().numpy(), OracleVecEnv.one_hot(g["view5"]))
    sym = mb.SymbolicObsWrapper(env).observation(obs)["image"]
    assert sym.dtype == torch.int64
    np.testing.assert_array_equal(sym.cpu().numpy(), g["symbolic"])
    np.testing.assert_array_equal(mb.RGBImgPartialObsWrapper(env).observation(obs)["image"].cpu().numpy(), g["rgb_partial"])
    np.testing.assert_array_equal(mb.RGBImgObsWrapper(env).observation(obs)["image"].cpu().numpy(), g["rgb_full"])
    np.testing.assert_array_equal(mb.FullyObsWrapper(env).observation(obs)["image"].cpu().numpy(), g["full_obs"])
    if "flat" in g:
        np.testing.assert_array_equal(mb.FlatObsWrapper(env).observation(obs).cpu().numpy(), g["flat"])


@pytest.mark.gpu
@pytest.mark.parametrize("env_id", ["MiniGrid-DoorKey-8x8-v0", "MiniGrid-MultiRoom-N6-v0", "MiniGrid-Dynamic-Obstacles-8x8-v0"])
def test_device_wrappers_follow_a_rollout(env_id):
    """Wrappers stacked on a stepping env (reset / step through the wrapper chain) against the oracle, 4133 envs."""
    import torch

    import minigrid_b200 as mb

    n = 4133
    orc = OracleVecEnv(env_id, n, n_threads=0)
    env = mb.OneHotPartialObsWrapper(mb.ViewSizeWrapper(mb.MinigridVecEnv(env_id, n), agent_view_size=9))
    sym = mb.SymbolicObsWrapper(mb.MinigridVecEnv(env_id, n))
    obs, _ = env.reset(seed=3)
    s_obs, _ = sym.reset(seed=3)
    orc.reset(seed=3)
    np.testing.assert_array_equal(obs["image"].cpu().numpy(), OracleVecEnv.one_hot(orc.gen_obs_view(9)))
    rng = np.random.default_rng(0)
    for t in range(60):
        a = rng.integers(0, 7, n).astype(np.int32)
        o = orc.step(a)
        obs, r, te, tr, _ = env.step(torch.as_tensor(a, device="cuda"))
        s_obs, *_ = sym.step(torch.as_tensor(a, device="cuda"))
        np.testing.assert_array_equal(obs["image"].cpu().numpy(), OracleVecEnv.one_hot(orc.gen_obs_view(9)), err_msg=f"t={t}")
        np.testing.assert_array_equal(s_obs["image"].cpu().numpy(), orc.symbolic_obs(), err_msg=f"symbolic t={t}")
        assert r.cpu().numpy().tobytes() == o[2].tobytes()
