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
