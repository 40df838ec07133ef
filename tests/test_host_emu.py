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
                                    "MiniGrid-Empty-5x5-v0", "MiniGrid-LavaCrossingS11N5-v0"])
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
