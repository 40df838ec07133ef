"""The C oracle (oracle/mg_oracle.c) against fixtures produced by the Python reference itself
(oracle/gen_golden.py) and against the reference's known-answer vectors. CPU only."""
import os

import numpy as np
import pytest
from conftest import golden_files, load_golden

from oracle.oracle import ENV_SPECS, OracleVecEnv


@pytest.mark.parametrize("path", golden_files("rollout") + golden_files("rollout_samestep"), ids=os.path.basename)
def test_rollout_matches_reference_fixture(path):
    g = load_golden(path)
    n = g["actions"].shape[1]
    o = OracleVecEnv(g["env_id"], n, autoreset=g["mode"])
    assert (o.width, o.height, o.max_steps, o.see_through) == (g["width"], g["height"], g["max_steps"], g["see_through"])
    obs, d = o.reset(seed=g["seed"])
    np.testing.assert_array_equal(obs, g["obs0"])
    np.testing.assert_array_equal(d, g["dir0"])
    st = o.get_state()
    np.testing.assert_array_equal(st["grid"], g["grid0"])
    np.testing.assert_array_equal(st["agent"], g["agent0"])
    np.testing.assert_array_equal(st["rng"], g["rng0"])
    np.testing.assert_array_equal(o.full_obs(), g["full_obs0"])
    for t in range(g["actions"].shape[0]):
        obs, d, r, te, tr = o.step(g["actions"][t])
        np.testing.assert_array_equal(obs, g["obs"][t], err_msg=f"obs t={t}")
        np.testing.assert_array_equal(d, g["dir"][t])
        assert r.tobytes() == g["reward"][t].tobytes(), f"reward bits t={t}"
        np.testing.assert_array_equal(te, g["terminated"][t])
        np.testing.assert_array_equal(tr, g["truncated"][t])
    st = o.get_state()
    for k in ("grid", "agent", "rng", "pending"):
        np.testing.assert_array_equal(st[k], g[k], err_msg=k)
    np.testing.assert_array_equal(o.full_obs(), g["full_obs"])


def make_oracle_wrapped(g):
    """make_env for parity.check_rollout_fixture: the oracle with the fixture's reward wrappers."""
    def make(env_id, n, mode):
        o = OracleVecEnv(env_id, n, autoreset=mode)
        o.set_no_death([str(t) for t in g["no_death"]], float(g["death_cost"]))
        o.set_bonus(str(g["bonus"]) or None)
        return o
    return make


@pytest.mark.parametrize("path", golden_files("rewardwrap"), ids=os.path.basename)
def test_reward_wrappers_match_reference_fixture(path):
    """NoDeath / ActionBonus / PositionBonus (wrappers.py:68-184, 809-882) applied by the reference itself (oracle/gen_golden.py)."""
    import parity

    g = load_golden(path)
    parity.check_rollout_fixture(make_oracle_wrapped(g), g)


@pytest.mark.parametrize("path", golden_files("inject"), ids=os.path.basename)
def test_injected_states_match_reference_fixture(path):
    g = load_golden(path)
    n = g["actions"].shape[1]
    o = OracleVecEnv(g["env_id"], n, autoreset="disabled")
    o.set_state(grid=g["grid0"], agent=g["agent0"])
    obs, _ = o.gen_obs()
    np.testing.assert_array_equal(obs, g["obs0"])
    for t in range(g["actions"].shape[0]):
        obs, d, r, te, tr = o.step(g["actions"][t])
        np.testing.assert_array_equal(obs, g["obs"][t], err_msg=f"obs t={t}")
        np.testing.assert_array_equal(d, g["dir"][t])
        assert r.tobytes() == g["reward"][t].tobytes()
        np.testing.assert_array_equal(te, g["terminated"][t])
        np.testing.assert_array_equal(tr, g["truncated"][t])
    st = o.get_state()
    np.testing.assert_array_equal(st["grid"], g["grid"])
    np.testing.assert_array_equal(st["agent"], g["agent"])


def test_reference_doctest_rng_kat():
    """minigrid/wrappers.py:26-41 (ReseedWrapper doctest): Empty-5x5 reset(seed) draws nothing and the
    next 10 np_random.integers(10) are these."""
    kat = {123: [0, 6, 5, 0, 9, 2, 2, 1, 3, 1], 0: [8, 6, 5, 2, 3, 0, 0, 0, 1, 8], 1: [4, 5, 7, 9, 0, 1, 8, 9, 2, 3]}
    for seed, want in kat.items():
        o = OracleVecEnv("MiniGrid-Empty-5x5-v0", 1)
        o.reset(seed=seed)
        assert [o.rng_integers(0, 0, 10) for _ in range(10)] == want


def test_reference_doctest_img_kat():
    """minigrid/wrappers.py:226-234: Empty-5x5 obs['image'][0,:,:] is seven grey walls."""
    o = OracleVecEnv("MiniGrid-Empty-5x5-v0", 1)
    obs, _ = o.reset(seed=0)
    np.testing.assert_array_equal(obs[0, 0], np.tile(np.array([2, 5, 0], np.uint8), (7, 1)))


def test_reference_no_death_kat():
    """tests/test_wrappers.py:364-380 / wrappers.py:818-830: LavaCrossingS9N1 seed 2, step(1), step(2)
    walks into lava: reward 0, terminated."""
    o = OracleVecEnv("MiniGrid-LavaCrossingS9N1-v0", 1)
    o.reset(seed=2)
    _, _, r, te, _ = o.step([1])
    assert not te[0]
    _, _, r, te, _ = o.step([2])
    assert te[0] and r[0] == 0.0


def test_truncation_exactly_at_max_steps():
    """tests/test_envs.py:160-177: truncated first at step == max_steps (action 4 = drop)."""
    o = OracleVecEnv(spec=("empty", 8, 8, 50, True, [0, 1, 1, 0]), num_envs=1, autoreset="disabled")
    o.reset(seed=0)
    for t in range(1, 51):
        _, _, _, te, tr = o.step([4])
        assert tr[0] == (t == 50) and not te[0]


def test_agent_sees_invariant():
    """tests/test_envs.py:133-154 restated on arrays: the goal is in the decoded image iff its view cell
    passes the visibility mask (here: type 8 present in obs <=> goal within view and seen)."""
    o = OracleVecEnv("MiniGrid-DoorKey-8x8-v0", 4)
    o.reset(seed=3)
    rng = np.random.default_rng(0)
    for _ in range(300):
        obs, *_ = o.step(rng.integers(0, 7, 4))
        assert set(np.unique(obs[..., 0])) <= {0, 1, 2, 4, 5, 8}


def test_invalid_action_raises():
    o = OracleVecEnv("MiniGrid-Empty-5x5-v0", 2)
    o.reset(seed=0)
    with pytest.raises(ValueError):
        o.step([0, 7])


def test_spec_table_is_consistent():
    for env_id, (kind, w, h, ms, st, prm) in ENV_SPECS.items():
        assert kind in ("empty", "doorkey", "crossing", "fourrooms", "lavagap", "distshift", "multiroom", "lockedroom", "playground", "gotodoor", "fetch", "redbluedoors", "gotoobject", "putnear", "memory", "dynobstacles", "roomgrid") and w >= 3 and h >= 3 and ms > 0
