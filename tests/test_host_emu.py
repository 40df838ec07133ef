"""CPU check of the engine's device arithmetic: minigrid_b200/csrc/*.cuh compiled by g++ (tests/host_emu)
against the reference-generated fixtures and the oracle. Catches layout / bit-trick bugs before GPU time."""
import os
import sys

import numpy as np
import pytest
from conftest import golden_files, load_golden

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emu"))
import parity  # noqa: E402
from emu import EmuVecEnv  # noqa: E402

from oracle.oracle import ENV_SPECS, OracleVecEnv  # noqa: E402


def make_emu(env_id, n, mode, layout=-1):
    return EmuVecEnv(ENV_SPECS[env_id], n, autoreset=mode, layout=layout)


@pytest.mark.parametrize("path", golden_files("rollout") + golden_files("rollout_samestep"), ids=os.path.basename)
def test_emu_rollout_fixture(path):
    parity.check_rollout_fixture(make_emu, load_golden(path))


@pytest.mark.parametrize("path", golden_files("inject"), ids=os.path.basename)
def test_emu_inject_fixture(path):
    parity.check_inject_fixture(make_emu, load_golden(path))


@pytest.mark.parametrize("env_id", ["MiniGrid-DoorKey-8x8-v0", "MiniGrid-FourRooms-v0", "MiniGrid-LavaCrossingS9N1-v0",
                                    "MiniGrid-Empty-5x5-v0", "MiniGrid-LavaCrossingS11N5-v0",
                                    "MiniGrid-Dynamic-Obstacles-5x5-v0", "MiniGrid-Dynamic-Obstacles-Random-6x6-v0",
                                    "MiniGrid-Dynamic-Obstacles-8x8-v0", "MiniGrid-Dynamic-Obstacles-16x16-v0",
                                    "MiniGrid-RedBlueDoors-8x8-v0", "MiniGrid-PutNear-6x6-N2-v0", "MiniGrid-LockedRoom-v0"])
@pytest.mark.parametrize("mode,n", [("next_step", 96), ("same_step", 45)])
def test_emu_lockstep_vs_oracle(env_id, mode, n):
    emu = make_emu(env_id, n, mode)
    orc = OracleVecEnv(env_id, n, autoreset=mode)
    parity.check_lockstep_vs_oracle(emu, orc, 300, seed=4242, check_state_every=100)


@pytest.mark.parametrize("layout", [0, 1], ids=["tiled", "window"])
@pytest.mark.parametrize("path", golden_files("rollout")[:6] + golden_files("inject"), ids=os.path.basename)
def test_emu_fixture_in_both_layouts(path, layout):
    """Every grid size through both HBM layouts (mg_create picks one by size; the other must agree)."""
    g = load_golden(path)
    mk = lambda env_id, n, mode: make_emu(env_id, n, mode, layout)  # noqa: E731
    if "rng0" in g:
        parity.check_rollout_fixture(mk, g)
    else:
        parity.check_inject_fixture(mk, g)


@pytest.mark.parametrize("env_id,layout,scalar,mode", [
    ("MiniGrid-DoorKey-8x8-v0", 0, False, "next_step"), ("MiniGrid-DoorKey-8x8-v0", 1, False, "same_step"),
    ("MiniGrid-FourRooms-v0", 1, False, "next_step"), ("MiniGrid-FourRooms-v0", 0, True, "same_step"),
    ("MiniGrid-Empty-5x5-v0", 0, False, "same_step"), ("MiniGrid-Empty-5x5-v0", 1, True, "next_step"),
    ("MiniGrid-Fetch-8x8-N3-v0", 0, False, "next_step"), ("MiniGrid-GoToDoor-5x5-v0", 1, False, "same_step"),
    ("MiniGrid-Dynamic-Obstacles-6x6-v0", 0, False, "same_step"), ("MiniGrid-Dynamic-Obstacles-6x6-v0", 0, True, "next_step"),
    ("MiniGrid-ObstructedMaze-1Dlh-v0", 0, False, "next_step"), ("MiniGrid-ObstructedMaze-1Dlh-v0", 1, True, "next_step")])  # cell code 13
def test_packed_host_format_expands_to_the_same_arrays(env_id, layout, scalar, mode, monkeypatch):
    """MG_HOST_PACKED: K1's 52-byte records (pack_codes, device header) + the product's host expander == the oracle's
    obs / dir / reward / flags, bit for bit, with both expander code paths."""
    import subprocess

    code = f"""
import sys, os
sys.path.insert(0, {os.path.dirname(os.path.abspath(__file__))!r}); sys.path.insert(0, {os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emu")!r})
import numpy as np
from emu import EmuVecEnv
from oracle.oracle import ENV_SPECS, OracleVecEnv
n = 77
emu = EmuVecEnv(ENV_SPECS[{env_id!r}], n, autoreset={mode!r}, layout={layout})
orc = OracleVecEnv({env_id!r}, n, autoreset={mode!r})
emu.reset(seed=5); orc.reset(seed=5)
rng = np.random.default_rng(0)
seen_reward = 0
for t in range(300):
    a = rng.integers(0, 7, n).astype(np.int32)
    if t % 3 == 0: a[:] = np.where(rng.random(n) < 0.6, 2, a)  # mostly forward: reach goals
    e = emu.step_packed(a); o = orc.step(a)
    assert np.array_equal(e[0], o[0]), t
    assert np.array_equal(e[1], o[1]) and e[2].tobytes() == o[2].tobytes()
    assert np.array_equal(e[3], o[3]) and np.array_equal(e[4], o[4])
    seen_reward += int((o[2] != 0).sum())
print("rewards", seen_reward)
"""
    env = dict(os.environ)
    if scalar:
        env["MINIGRID_B200_EXPAND_SCALAR"] = "1"  # read once per process: hence the subprocess
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stderr[-3000:]
    if env_id == "MiniGrid-Empty-5x5-v0":
        assert int(r.stdout.split()[-1]) > 0  # the reward path was exercised


@pytest.mark.parametrize("layout", [0, 1], ids=["tiled", "window"])
@pytest.mark.parametrize("env_id", ["MiniGrid-Unlock-v0", "MiniGrid-UnlockPickup-v0", "MiniGrid-BlockedUnlockPickup-v0",
                                    "MiniGrid-KeyCorridorS3R3-v0", "MiniGrid-KeyCorridorS6R3-v0"])
@pytest.mark.parametrize("mode", ["next_step", "same_step"])
def test_emu_roomgrid_post_filters(env_id, layout, mode):
    """The success branches of the RoomGrid step post-filters (unlock.py:88-96, `carrying == self.obj`) through the device
    headers, against the oracle (itself pinned on these branches against the live reference, tests/test_oracle_next.py)."""
    n = 40
    emu = make_emu(env_id, n, mode, layout)
    orc = OracleVecEnv(env_id, n, autoreset=mode)
    parity.check_lockstep_vs_oracle(emu, orc, 20, seed=77)
    agent, act = parity.roomgrid_inject_targets(orc)
    emu.set_state(agent=agent)
    orc.set_state(agent=agent)
    rewarded = 0
    for a in (np.full(n, act), np.full(n, act), np.full(n, 2), np.full(n, act)):
        e, o = emu.step(a), orc.step(a)
        for x, y, name in zip(e, o, ["obs", "dir", "reward", "terminated", "truncated"]):
            if name == "reward":
                assert np.asarray(x).tobytes() == np.asarray(y).tobytes()
            else:
                np.testing.assert_array_equal(np.asarray(x).astype(np.asarray(y).dtype), y, err_msg=name)
        rewarded += int((o[2] > 0).sum())
    assert rewarded >= n // 2
    es, os_ = emu.get_state(), orc.get_state()
    for k in ("grid", "agent", "rng", "pending"):
        np.testing.assert_array_equal(es[k], os_[k], err_msg=k)


@pytest.mark.parametrize("layout", [0, 1], ids=["tiled", "window"])
@pytest.mark.parametrize("env_id", ["MiniGrid-ObstructedMaze-1Dlh-v0", "MiniGrid-ObstructedMaze-Full-v1"])
def test_emu_boxes_hide_keys(env_id, layout):
    """Box.contains (world_object.py:273-293): toggling a box of ObstructedMaze leaves its key; a box that is picked up
    and dropped again keeps it. Device headers (cell code 13) against the oracle."""
    n = 24
    emu = make_emu(env_id, n, "next_step", layout)
    orc = OracleVecEnv(env_id, n)
    parity.check_lockstep_vs_oracle(emu, orc, 5, seed=11)
    agent, moved = parity.face_first_cell_of_type(orc, 7)
    assert moved.sum() >= n // 2
    emu.set_state(agent=agent)
    orc.set_state(agent=agent)
    half = np.arange(n) % 2 == 0
    # even envs: toggle (box -> key), pickup (the key); odd envs: pickup (the box), turn, drop, toggle, pickup
    script = [np.where(half, 5, 3), np.where(half, 3, 0), np.where(half, 6, 4), np.where(half, 6, 5), np.where(half, 4, 3), np.full(n, 2)]
    for a in script:
        e, o = emu.step(a), orc.step(a)
        for x, y, name in zip(e, o, ["obs", "dir", "reward", "terminated", "truncated"]):
            np.testing.assert_array_equal(np.asarray(x).astype(np.asarray(y).dtype), y, err_msg=name)
    es, os_ = emu.get_state(), orc.get_state()
    for k in ("grid", "agent", "rng", "pending"):
        np.testing.assert_array_equal(es[k], os_[k], err_msg=k)
    assert (os_["agent"][moved, 3] == 5).sum() >= moved.sum() // 4  # keys did come out of boxes


def test_expand_pool_matches_the_single_threaded_expander():
    """The host pool (slices handed out through generation-tagged counters, the caller helping) against the plain
    single-threaded expander on random records: many back-to-back jobs of changing size and thread count. Host code of
    the product library, no GPU needed."""
    import ctypes as C

    from minigrid_b200 import _build

    L = C.CDLL(_build.LIB_PATH)
    p = C.c_void_p
    L.mg_expand_packed.argtypes = [p, C.c_int64, C.c_int32, p, p, p, p, p]
    L.mg_expand_packed_mt.argtypes = [p, C.c_int64, C.c_int32, p, p, p, p, p, C.c_int]
    rng = np.random.default_rng(3)
    ptr = lambda a: a.ctypes.data_as(p)
    for it in range(60):
        n = int(rng.integers(1, 70000))
        packed = rng.integers(0, 256, (n, 52), dtype=np.uint8)
        packed[:, 0:49] = rng.integers(0, 14, (n, 49)) | (rng.integers(0, 6, (n, 49)) << 4)  # plausible cell codes
        packed[:, 49] &= 0x1F          # step_count < 2^16
        packed[:, 50:52] = rng.integers(0, 256, (n, 2)); packed[:, 51] &= 0x01  # step_count <= 511
        outs = []
        for mt in (0, int(rng.integers(1, 9))):
            obs = np.zeros((n, 147), np.uint8); d = np.zeros(n, np.int32); r = np.zeros(n, np.float64)
            te = np.zeros(n, np.uint8); tr = np.zeros(n, np.uint8)
            if mt == 0:
                assert L.mg_expand_packed(ptr(packed), n, 640, ptr(obs), ptr(d), ptr(r), ptr(te), ptr(tr)) == 0
            else:
                assert L.mg_expand_packed_mt(ptr(packed), n, 640, ptr(obs), ptr(d), ptr(r), ptr(te), ptr(tr), mt) == 0
            outs.append((obs, d, r, te, tr))
        for a, b in zip(*outs):
            assert a.tobytes() == b.tobytes(), f"job {it} n={n}"
