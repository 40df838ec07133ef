"""Shared parity drivers: replay a golden fixture / lockstep against the oracle on any engine object that
exposes reset/step/gen_obs/full_obs/get_state/set_state (the CUDA engine through the C-ABI, or the host
emulation of its device headers)."""
import numpy as np


def check_rollout_fixture(make_env, g):
    n = g["actions"].shape[1]
    e = make_env(g["env_id"], n, g["mode"])
    obs, d = e.reset(seed=g["seed"])
    np.testing.assert_array_equal(np.asarray(obs), g["obs0"])
    np.testing.assert_array_equal(np.asarray(d), g["dir0"])
    st = e.get_state()
    np.testing.assert_array_equal(st["grid"], g["grid0"])
    np.testing.assert_array_equal(st["agent"], g["agent0"])
    np.testing.assert_array_equal(st["rng"], g["rng0"])
    np.testing.assert_array_equal(e.full_obs(), g["full_obs0"])
    for t in range(g["actions"].shape[0]):
        obs, d, r, te, tr = e.step(g["actions"][t])
        np.testing.assert_array_equal(np.asarray(obs), g["obs"][t], err_msg=f"obs t={t}")
        np.testing.assert_array_equal(np.asarray(d), g["dir"][t], err_msg=f"dir t={t}")
        assert np.asarray(r, np.float64).tobytes() == g["reward"][t].tobytes(), f"reward bits t={t}"
        np.testing.assert_array_equal(np.asarray(te, bool), g["terminated"][t], err_msg=f"terminated t={t}")
        np.testing.assert_array_equal(np.asarray(tr, bool), g["truncated"][t], err_msg=f"truncated t={t}")
    st = e.get_state()
    for k in ("grid", "agent", "rng", "pending"):
        np.testing.assert_array_equal(st[k], g[k], err_msg=k)
    np.testing.assert_array_equal(e.full_obs(), g["full_obs"])


def check_inject_fixture(make_env, g):
    n = g["actions"].shape[1]
    e = make_env(g["env_id"], n, "disabled")
    e.reset(seed=0)
    e.set_state(grid=g["grid0"], agent=g["agent0"])
    obs, _ = e.gen_obs()
    np.testing.assert_array_equal(np.asarray(obs), g["obs0"])
    for t in range(g["actions"].shape[0]):
        obs, d, r, te, tr = e.step(g["actions"][t])
        np.testing.assert_array_equal(np.asarray(obs), g["obs"][t], err_msg=f"obs t={t}")
        np.testing.assert_array_equal(np.asarray(d), g["dir"][t])
        assert np.asarray(r, np.float64).tobytes() == g["reward"][t].tobytes()
        np.testing.assert_array_equal(np.asarray(te, bool), g["terminated"][t])
        np.testing.assert_array_equal(np.asarray(tr, bool), g["truncated"][t])
    st = e.get_state()
    np.testing.assert_array_equal(st["grid"], g["grid"])
    np.testing.assert_array_equal(st["agent"], g["agent"])


def check_lockstep_vs_oracle(engine, oracle, n_steps, seed, action_seed=1234, check_state_every=0):
    """engine and oracle are already constructed with the same spec/mode/num_envs."""
    eo, ed = engine.reset(seed=seed)
    oo, od = oracle.reset(seed=seed)
    np.testing.assert_array_equal(np.asarray(eo), oo)
    np.testing.assert_array_equal(np.asarray(ed), od)
    rng = np.random.default_rng(action_seed)
    n = oracle.num_envs
    for t in range(n_steps):
        a = rng.integers(0, 7, n).astype(np.int32)
        e = engine.step(a)
        o = oracle.step(a)
        for x, y, name in zip(e, o, ["obs", "dir", "reward", "terminated", "truncated"]):
            x = np.asarray(x); y = np.asarray(y)
            if name == "reward":
                assert x.astype(np.float64).tobytes() == y.tobytes(), f"reward bits t={t}"
            else:
                np.testing.assert_array_equal(x.astype(y.dtype), y, err_msg=f"{name} t={t}")
        if check_state_every and (t + 1) % check_state_every == 0:
            es, os_ = engine.get_state(), oracle.get_state()
            for k in ("grid", "agent", "rng", "pending"):
                np.testing.assert_array_equal(es[k], os_[k], err_msg=f"{k} t={t}")
    es, os_ = engine.get_state(), oracle.get_state()
    for k in ("grid", "agent", "rng", "pending"):
        np.testing.assert_array_equal(es[k], os_[k], err_msg=k)
    np.testing.assert_array_equal(engine.full_obs(), oracle.full_obs())


def roomgrid_inject_targets(orc):
    """Agents placed next to their env's target (Unlock: in front of the locked door with its key; the pickup variants:
    facing the box / ball), from the oracle's state. Returns (agent array, action)."""
    st = orc.get_state()
    grid, agent = st["grid"], st["agent"].copy()
    unlock = orc.params[0] == 0
    want = 4 if unlock else (6 if orc.params[0] == 3 else 7)  # door | ball (KeyCorridor) | box
    for i in range(orc.num_envs):
        xs, ys = np.nonzero((grid[i, :, :, 0] == want) & ((grid[i, :, :, 2] == 2) if unlock else True))
        tx, ty = int(xs[0]), int(ys[0])
        for d, (dx, dy) in enumerate([(1, 0), (0, 1), (-1, 0), (0, -1)]):
            ax, ay = tx - dx, ty - dy
            if 0 < ax < orc.width - 1 and 0 < ay < orc.height - 1 and grid[i, ax, ay, 0] in ((1,) if unlock else (1, 4)):
                agent[i, :3] = (ax, ay, d)
                agent[i, 3:5] = (5, grid[i, tx, ty, 1]) if unlock else (-1, 0)
                break
    return agent, (5 if unlock else 3)


def face_first_cell_of_type(orc, want_type):
    """Agent records that put every agent next to (and facing) the first cell of the given type of its env, from the
    oracle's state; envs without such a cell or without a free neighbour keep their record."""
    st = orc.get_state()
    grid, agent = st["grid"], st["agent"].copy()
    moved = np.zeros(orc.num_envs, bool)
    for i in range(orc.num_envs):
        xs, ys = np.nonzero(grid[i, :, :, 0] == want_type)
        if len(xs) == 0:
            continue
        tx, ty = int(xs[0]), int(ys[0])
        for d, (dx, dy) in enumerate([(1, 0), (0, 1), (-1, 0), (0, -1)]):
            ax, ay = tx - dx, ty - dy
            if 0 < ax < orc.width - 1 and 0 < ay < orc.height - 1 and grid[i, ax, ay, 0] == 1:
                agent[i, :3] = (ax, ay, d)
                agent[i, 3:5] = (-1, 0)
                moved[i] = True
                break
    return agent, moved
