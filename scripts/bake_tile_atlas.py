"""Bakes minigrid_b200/data/tile_atlas.npz: every 8 x 8 tile the reference's Grid.render_tile (grid.py:145-198) can draw,
rendered BY THE UNMODIFIED REFERENCE (build container only), indexed by (cell code, agent overlay, highlight). The
device-side RGB wrappers (mg_wrappers.cu: k_rgb_partial, k_rgb_full) only copy tile rows, so their output is the
reference's pixels by construction. Run: python scripts/bake_tile_atlas.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

ref_loader.load()
from minigrid.core.grid import Grid  # noqa: E402
from minigrid.core.world_object import WorldObj  # noqa: E402


def decode_code(code):
    """cell code (minigrid_b200/csrc/mg_common.cuh) -> (type, colour, state) or None for 'no object'"""
    t4, colour = code & 15, (code >> 4) & 7
    if t4 in (0, 1):
        return None
    if colour > 5 or t4 == 10 or t4 > 12:
        return "invalid"
    if t4 == 13:
        return 7, 5, 0  # a grey box with a key inside (ObstructedMaze): rendered like any grey box
    if t4 == 11:
        return 4, colour, 1
    if t4 == 12:
        return 4, colour, 2
    return t4, colour, 0


tiles, index = [], np.zeros((128, 5, 2), np.uint16)
cache = {}
for code in range(128):
    d = decode_code(code)
    for agent in range(5):
        for hl in range(2):
            if d == "invalid":
                index[code, agent, hl] = 0  # filled below with the empty tile of the same overlay
                continue
            obj = None if d is None else WorldObj.decode(*d)
            key = (d, agent, hl)
            if key not in cache:
                img = Grid.render_tile(obj, agent_dir=None if agent == 0 else agent - 1, highlight=bool(hl), tile_size=8)
                assert img.shape == (8, 8, 3)
                # render_tile returns the float64 means of downsample(); Grid.render stores them into a uint8 image
                # (grid.py:217-241: `img[ymin:ymax, xmin:xmax, :] = tile_img`), i.e. the same cast as here
                px = np.zeros((8, 8, 3), np.uint8)
                px[:, :, :] = img
                cache[key] = len(tiles)
                tiles.append(px)
            index[code, agent, hl] = cache[key]
for code in range(128):
    if decode_code(code) == "invalid":
        index[code] = index[1]
tiles = np.stack(tiles)
out = os.path.join(ROOT, "minigrid_b200", "data", "tile_atlas.npz")
np.savez_compressed(out, tiles=tiles, index=index, tile_size=np.int32(8))
print("tiles", tiles.shape, "->", out, os.path.getsize(out), "bytes")
