"""The C-ABI shared library loads without a GPU and exports every symbol include/minigrid_b200.h declares;
error paths that need no device behave (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "minigrid_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mg_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from minigrid_b200 import _lib

    L = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 18
    for name in syms:
        assert hasattr(L, name), f"{name} declared in include/minigrid_b200.h but not exported"
    assert set(_lib.EXPORTS) <= set(syms)


def test_create_rejects_bad_arguments_without_device_work():
    from minigrid_b200 import _lib

    L = _lib.load()
    h = C.c_void_p()
    assert L.mg_create(99, 8, 8, 100, 0, None, 0, 4, 0, 0, C.byref(h)) == -1
    assert b"kind" in L.mg_last_error()
    assert L.mg_create(0, 40, 8, 100, 0, None, 0, 4, 0, 0, C.byref(h)) == -1
    assert L.mg_create(2, 8, 8, 100, 0, None, 0, 4, 0, 0, C.byref(h)) == -1  # crossing needs odd sizes
    assert L.mg_create(0, 8, 8, 100, 0, None, 0, 0, 0, 0, C.byref(h)) == -1
    assert not h.value


def test_product_fails_loudly_without_cuda():
    import torch

    import minigrid_b200

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(minigrid_b200.MinigridB200Error):
        minigrid_b200.MinigridVecEnv("MiniGrid-Empty-5x5-v0", 4)
    from minigrid_b200 import _lib

    L = _lib.load()
    h = C.c_void_p()
    assert L.mg_create(0, 8, 8, 100, 1, None, 0, 4, 0, 0, C.byref(h)) == -4  # MG_ERR_NO_DEVICE, no CPU fallback


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "minigrid_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src, f"{fn} mentions the oracle"


def test_spec_tables_agree():
    from minigrid_b200 import specs
    from oracle.oracle import ENV_SPECS

    kinds = ["empty", "doorkey", "crossing", "fourrooms", "lavagap", "distshift", "multiroom", "lockedroom", "playground", "gotodoor", "fetch", "redbluedoors", "gotoobject", "putnear", "memory", "dynobstacles", "roomgrid"]
    for env_id, (kind, w, h, ms, st, prm) in ENV_SPECS.items():
        s = specs.get(env_id)
        assert (kinds[s.kind], s.width, s.height, s.max_steps, s.see_through_walls) == (kind, w, h, ms, st), env_id
        assert list(s.params) == list(prm), env_id


def test_shard_ranges_tile_the_batch():
    from minigrid_b200 import shard_range

    for total in (1, 7, 64, 1000, 2097152):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
