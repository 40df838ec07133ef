"""Live check of the C oracle against the unmodified Python reference (build container only: skipped
where /root/reference is absent, e.g. on the GPU box)."""
import numpy as np
import pytest

from oracle import ref_loader
from oracle.oracle import ENV_SPECS, OracleVecEnv

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


@pytest.mark.parametrize("env_id", list(ENV_SPECS))
@pytest.mark.parametrize("mode", ["next_step", "same_step"])
def test_lockstep_rollout(env_id, mode):
    n, t_steps = 6, 250
    ref = ref_loader.ReferenceVecEnv(env_id, n, autoreset=mode)
    orc = OracleVecEnv(env_id, n, autoreset=mode)
    e0 = ref.envs[0]
    assert (orc.width, orc.height, orc.max_steps, orc.see_through) == (e0.width, e0.height, e0.max_steps, e0.see_through_walls)
    ro, rd = ref.reset(seed=1000)
    oo, od = orc.reset(seed=1000)
    np.testing.assert_array_equal(ro, oo)
    np.testing.assert_array_equal(rd, od)
    rng = np.random.default_rng(77)
    for t in range(t_steps):
        a = rng.integers(0, 7, n)
        r = ref.step(a)
        q = orc.step(a)
        for x, y, name in zip(r, q, ["obs", "dir", "reward", "terminated", "truncated"]):
            np.testing.assert_array_equal(np.asarray(x), np.asarray(y), err_msg=f"{name} t={t}")
    rs, os_ = ref.get_state(), orc.get_state()
    for k in rs:
        np.testing.assert_array_equal(rs[k], os_[k], err_msg=k)
    np.testing.assert_array_equal(ref.full_obs(), orc.full_obs())


def test_spec_table_matches_registry():
    gym, _ = ref_loader.load()
    for env_id, (kind, w, h, ms, st, prm) in ENV_SPECS.items():
        e = gym.make(env_id).unwrapped
        assert (e.width, e.height, e.max_steps, e.see_through_walls) == (w, h, ms, st), env_id


@pytest.mark.parametrize("base_id,kind,size,max_steps,see_through,params", [
    ("MiniGrid-Empty-5x5-v0", "empty", 26, 4 * 26 * 26, True, [1, 0, 0, 0]),  # agent_start_pos=None: random start
    ("MiniGrid-DoorKey-5x5-v0", "doorkey", 26, 10 * 26 * 26, False, []),
    ("MiniGrid-DoorKey-5x5-v0", "doorkey", 6, 10 * 6 * 6, False, []),
])
def test_lockstep_rollout_at_unregistered_sizes(base_id, kind, size, max_steps, see_through, params):
    """The engine's size limit (26) and a small DoorKey that no id registers: the reference classes take `size`."""
    n, t_steps = 4, 200
    kwargs = {"size": size}
    if kind == "empty":
        kwargs["agent_start_pos"] = None
    ref = ref_loader.ReferenceVecEnv(base_id, n, **kwargs)
    orc = OracleVecEnv(None, n, spec=(kind, size, size, max_steps, see_through, params))
    e0 = ref.envs[0]
    assert (e0.width, e0.height, e0.max_steps, e0.see_through_walls) == (size, size, max_steps, see_through)
    ro, rd = ref.reset(seed=5)
    oo, od = orc.reset(seed=5)
    np.testing.assert_array_equal(ro, oo)
    np.testing.assert_array_equal(rd, od)
    rng = np.random.default_rng(6)
    for t in range(t_steps):
        a = rng.integers(0, 7, n)
        for x, y, name in zip(ref.step(a), orc.step(a), ["obs", "dir", "reward", "terminated", "truncated"]):
            np.testing.assert_array_equal(np.asarray(x), np.asarray(y), err_msg=f"{name} t={t}")
    rs, os_ = ref.get_state(), orc.get_state()
    for k in rs:
        np.testing.assert_array_equal(rs[k], os_[k], err_msg=k)


@pytest.mark.parametrize("env_id", ["MiniGrid-DoorKey-8x8-v0", "MiniGrid-FourRooms-v0", "MiniGrid-Empty-5x5-v0", "MiniGrid-MultiRoom-N6-v0",
                                    "MiniGrid-Playground-v0", "MiniGrid-Dynamic-Obstacles-6x6-v0"])
def test_observation_wrappers_against_live_reference(env_id):
    """ViewSizeWrapper (V = 3, 5, 9, 11), SymbolicObsWrapper and OneHotPartialObsWrapper restated in the
    oracle (gen_obs_view, symbolic_obs, one_hot) against the reference's own wrapper classes on the same states."""
    n = 6
    ref = ref_loader.ReferenceVecEnv(env_id, n)
    orc = OracleVecEnv(env_id, n)
    ref.reset(seed=31)
    orc.reset(seed=31)
    rng = np.random.default_rng(5)
    for t in range(120):
        a = rng.integers(0, 7, n)
        r = ref.step(a)
        q = orc.step(a)
        np.testing.assert_array_equal(r[0], q[0])
        if t % 6 == 0:
            for V in (3, 5, 7, 9, 11):
                np.testing.assert_array_equal(ref.view_obs(V), orc.gen_obs_view(V), err_msg=f"view {V} t={t}")
            sym = ref.symbolic_obs()
            assert sym.dtype == np.int64
            np.testing.assert_array_equal(sym, orc.symbolic_obs(), err_msg=f"symbolic t={t}")
            np.testing.assert_array_equal(ref.one_hot_obs(), OracleVecEnv.one_hot(q[0]), err_msg=f"one-hot t={t}")


# ---- SURVEY 8(f-4), second half: the reward wrappers (wrappers.py:68-184, 809-882) ----
def _wrap(no_death, bonus):
    ref_loader.load()
    from minigrid.wrappers import ActionBonus, NoDeath, PositionBonus

    def w(e):
        if no_death:
            e = NoDeath(e, no_death_types=no_death, death_cost=-1.5)
        if bonus == "action":
            e = ActionBonus(e)
        elif bonus == "position":
            e = PositionBonus(e)
        return e
    return w


@pytest.mark.parametrize("env_id,no_death,bonus", [
    ("MiniGrid-LavaCrossingS9N1-v0", ("lava",), None),
    ("MiniGrid-LavaCrossingS9N3-v0", ("lava",), "action"),
    ("MiniGrid-DistShift1-v0", ("lava",), "position"),
    ("MiniGrid-LavaGapS5-v0", ("lava", "wall"), None),
    ("MiniGrid-Dynamic-Obstacles-5x5-v0", ("ball",), None),
    ("MiniGrid-Dynamic-Obstacles-6x6-v0", ("ball",), "action"),
    ("MiniGrid-Empty-5x5-v0", (), "action"),
    ("MiniGrid-DoorKey-5x5-v0", (), "position"),
    ("MiniGrid-FourRooms-v0", (), "action"),
    ("MiniGrid-GoToDoor-5x5-v0", ("door",), "position"),
])
@pytest.mark.parametrize("mode", ["next_step", "same_step"])
def test_reward_wrappers_against_live_reference(env_id, no_death, bonus, mode):
    """NoDeath, ActionBonus and PositionBonus as a SyncVectorEnv of wrapped envs applies them (bonus outermost), restated
    in the oracle (wrapped_step): rewards bit for bit, and the episodes NoDeath keeps alive stay alive."""
    n, t_steps = 6, 400
    ref = ref_loader.ReferenceVecEnv(env_id, n, autoreset=mode, wrap=_wrap(no_death, bonus))
    orc = OracleVecEnv(env_id, n, autoreset=mode)
    orc.set_no_death(no_death, -1.5)
    orc.set_bonus(bonus)
    ref.reset(seed=2)
    orc.reset(seed=2)
    rng = np.random.default_rng(11)
    saved = 0
    for t in range(t_steps):
        # forward-heavy actions: walk into lava / obstacles often
        a = np.where(rng.random(n) < 0.5, 2, rng.integers(0, 7, n))
        r = ref.step(a)
        q = orc.step(a)
        for x, y, name in zip(r, q, ["obs", "dir", "reward", "terminated", "truncated"]):
            np.testing.assert_array_equal(np.asarray(x), np.asarray(y), err_msg=f"{name} t={t}")
        assert r[2].tobytes() == q[2].tobytes()
        saved += int(((r[2] < -0.4) & ~r[3]).sum())  # a negative reward on a live env: the death cost
    if no_death and "Empty" not in env_id and "GoToDoor" not in env_id:
        assert saved > 0, "NoDeath never triggered: the test does not cover it"


def test_reward_wrapper_known_answers():
    """The reference's own doctests: wrappers.py:81-93 (ActionBonus 1.0, 1.0), :137-145 (PositionBonus 1.0, 0.7071067811865475),
    :818-834 (NoDeath: LavaCrossingS9N1 seed 2 -> (-1.0, False); Dynamic-Obstacles-5x5 seed 2 -> (-2.0, False))."""
    o = OracleVecEnv("MiniGrid-Empty-5x5-v0", 1); o.set_bonus("action"); o.reset(seed=0)
    assert [float(o.step([1])[2][0]) for _ in range(2)] == [1.0, 1.0]
    o = OracleVecEnv("MiniGrid-Empty-5x5-v0", 1); o.set_bonus("position"); o.reset(seed=0)
    assert [float(o.step([1])[2][0]) for _ in range(2)] == [1.0, 0.7071067811865475]
    o = OracleVecEnv("MiniGrid-LavaCrossingS9N1-v0", 1); o.reset(seed=2); o.step([1])
    r = o.step([2]); assert (float(r[2][0]), bool(r[3][0])) == (0.0, True)
    o = OracleVecEnv("MiniGrid-LavaCrossingS9N1-v0", 1); o.set_no_death(("lava",), -1.0); o.reset(seed=2); o.step([1])
    r = o.step([2]); assert (float(r[2][0]), bool(r[3][0])) == (-1.0, False)
    o = OracleVecEnv("MiniGrid-Dynamic-Obstacles-5x5-v0", 1); o.reset(seed=2)
    r = o.step([2]); assert (float(r[2][0]), bool(r[3][0])) == (-1.0, True)
    o = OracleVecEnv("MiniGrid-Dynamic-Obstacles-5x5-v0", 1); o.set_no_death(("ball",), -1.0); o.reset(seed=2)
    r = o.step([2]); assert (float(r[2][0]), bool(r[3][0])) == (-2.0, False)


def test_dict_observation_space_wrapper_mission_indices():
    """minigrid_b200.wrappers.mission_to_indices against the reference's DictObservationSpaceWrapper (wrappers.py:428-554) on the
    constant mission strings of the registered ids (its doctest value included: LavaCrossingS11N5 -> [19, 31, 17, 36, 20, 38, ...])."""
    gym, _ = ref_loader.load()
    from minigrid.wrappers import DictObservationSpaceWrapper as RefDict

    from minigrid_b200 import specs
    from minigrid_b200.wrappers import MINIGRID_WORDS, mission_to_indices

    assert {w: i for i, w in enumerate(MINIGRID_WORDS)} == RefDict.get_minigrid_words()
    seen = 0
    for env_id in specs.all_ids() if hasattr(specs, "all_ids") else list(ENV_SPECS):
        mission = specs.get(env_id).mission
        if "{" in mission:
            continue
        e = RefDict(gym.make(env_id))
        obs, _ = e.reset(seed=0)
        assert obs["mission"] == mission_to_indices(mission), env_id
        seen += 1
    assert seen >= 20
    assert mission_to_indices("avoid the lava and get to the green goal square")[:10] == [19, 31, 17, 36, 20, 38, 31, 2, 15, 35]
